// GSScene.cpp -- see GSScene.h.  Mirrors the behaviour of src/GSScene.cpp:26-68,99-149 of the
// reference for Inria-format files: "element vertex <N>", then N fixed 62-float little-endian
// records.  Unlike the reference the header's property list is honoured (GSScene.h, PlyLayout):
// the canonical list takes the verbatim path, anything else is gathered by property name.
#include "GSScene.h"

#include <fcntl.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <thread>

#include "gsmath.h"

namespace {
constexpr size_t kRecordFloats = 62;  // VertexStorage: pos3 normal3 shs48 opacity scale3 rot4 (GSScene.cpp:17-24)
}

GSScene::GSScene(const std::string& file) : filename(file) {
    if (!std::filesystem::exists(filename)) throw std::runtime_error("File does not exist: " + filename);
}

void GSScene::loadPlyHeader(std::ifstream& plyFile) {
    if (!plyFile.is_open()) throw std::runtime_error("Could not open file: " + filename);
    std::string line;
    bool headerEnd = false;
    bool inVertexElement = false;
    while (std::getline(plyFile, line)) {
        std::istringstream iss(line);
        std::string token;
        iss >> token;
        if (token == "format") {
            iss >> header.format;
        } else if (token == "element") {
            iss >> token;
            inVertexElement = token == "vertex";
            if (token == "vertex") iss >> header.numVertices;
            else if (token == "face") iss >> header.numFaces;
        } else if (token == "property") {
            PlyProperty property;
            iss >> property.type >> property.name;
            if (property.type == "list") {  // "property list <count type> <item type> <name>"
                if (inVertexElement) throw std::runtime_error("list property in the vertex element of " + filename);
                std::string itemType;
                iss >> itemType >> property.name;
                property.type = "list";
            }
            (inVertexElement ? header.vertexProperties : header.faceProperties).push_back(property);
        } else if (token == "end_header") {
            headerEnd = true;
            break;
        }
    }
    if (!headerEnd) throw std::runtime_error("Could not find end of header");
    if (header.numVertices < 0) throw std::runtime_error("Negative vertex count in " + filename);
    resolveLayout();
}

namespace {
// canonical property names in record order
std::vector<std::string> canonicalNames() {
    std::vector<std::string> names = {"x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"};
    for (int k = 0; k < 45; k++) names.push_back("f_rest_" + std::to_string(k));
    names.push_back("opacity");
    for (int k = 0; k < 3; k++) names.push_back("scale_" + std::to_string(k));
    for (int k = 0; k < 4; k++) names.push_back("rot_" + std::to_string(k));
    return names;
}
int plyTypeSize(const std::string& t) {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "int32" || t == "uint32" || t == "float" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}
bool plyTypeIsFloat(const std::string& t) { return t == "float" || t == "float32"; }
bool plyTypeIsDouble(const std::string& t) { return t == "double" || t == "float64"; }
}  // namespace

void GSScene::resolveLayout() {
    layout = PlyLayout{};
    for (int k = 0; k < 62; k++) {
        layout.offset[k] = 4 * k;
        layout.is_double[k] = 0;
    }
    const auto& props = header.vertexProperties;
    if (!header.format.empty() && header.format != "binary_little_endian")
        throw std::runtime_error("unsupported PLY format '" + header.format + "' in " + filename + " (binary_little_endian only)");
    const std::vector<std::string> names = canonicalNames();
    if (props.empty()) return;  // no list at all: the reference's assumption is all there is
    bool same = props.size() == names.size();
    for (size_t k = 0; same && k < names.size(); k++) same = props[k].name == names[k] && plyTypeIsFloat(props[k].type);
    if (same) return;

    // gather by name
    layout.canonical = false;
    for (int k = 0; k < 62; k++) layout.offset[k] = -1;
    uint64_t at = 0;
    int restCount = 0;
    std::vector<int32_t> restOffset(45, -1);
    std::vector<uint8_t> restDouble(45, 0);
    for (const PlyProperty& pr : props) {
        const int size = plyTypeSize(pr.type);
        if (size == 0) throw std::runtime_error("unknown PLY property type '" + pr.type + "' in " + filename);
        if (at > 0x7fffffffull) throw std::runtime_error("vertex record too large in " + filename);
        const bool fl = plyTypeIsFloat(pr.type), db = plyTypeIsDouble(pr.type);
        if (pr.name.rfind("f_rest_", 0) == 0) {
            const int idx = std::atoi(pr.name.c_str() + 7);
            if (idx < 0 || idx >= 45 || !(fl || db)) throw std::runtime_error("bad SH property '" + pr.name + "' in " + filename);
            restOffset[idx] = static_cast<int32_t>(at);
            restDouble[idx] = db;
            restCount = std::max(restCount, idx + 1);
        } else {
            for (int k = 0; k < 62; k++) {
                if (k >= 9 && k < 54) continue;  // f_rest slots are resolved below
                if (names[k] == pr.name) {
                    if (!(fl || db)) throw std::runtime_error("PLY property '" + pr.name + "' is not a float in " + filename);
                    layout.offset[k] = static_cast<int32_t>(at);
                    layout.is_double[k] = db;
                }
            }
        }
        at += static_cast<uint64_t>(size);
    }
    layout.stride = at;
    // f_rest is channel-major with K = restCount / 3 coefficients per channel (K = 0, 3, 8, 15 for degree 0..3);
    // coefficient j of channel c goes to the canonical slot 15 c + j, the missing higher bands stay zero
    if (restCount % 3 != 0) throw std::runtime_error("f_rest count is not a multiple of 3 in " + filename);
    const int K = restCount / 3;
    if (K != 0 && K != 3 && K != 8 && K != 15) throw std::runtime_error("f_rest count matches no SH degree in " + filename);
    layout.shDegree = K == 0 ? 0 : (K == 3 ? 1 : (K == 8 ? 2 : 3));
    for (int c = 0; c < 3; c++)
        for (int j = 0; j < K; j++) {
            const int src = c * K + j;
            if (restOffset[src] < 0) throw std::runtime_error("missing f_rest_" + std::to_string(src) + " in " + filename);
            layout.offset[9 + 15 * c + j] = restOffset[src];
            layout.is_double[9 + 15 * c + j] = restDouble[src];
        }
    for (int k : {0, 1, 2, 6, 7, 8, 54, 55, 56, 57, 58, 59, 60, 61})
        if (layout.offset[k] < 0) throw std::runtime_error("missing PLY property '" + names[k] + "' in " + filename);
}

void GSScene::activateRecords(const float* records, uint64_t n, Vertex* out, unsigned threads) {
    auto work = [&](uint64_t begin, uint64_t end) {
        for (uint64_t i = begin; i < end; i++) {
            const float* s = records + i * kRecordFloats;
            const float* shs = s + 6;
            Vertex& v = out[i];
            v.position[0] = s[0];
            v.position[1] = s[1];
            v.position[2] = s[2];
            v.position[3] = 1.0f;
            v.scale_opacity[0] = std::exp(s[55]);
            v.scale_opacity[1] = std::exp(s[56]);
            v.scale_opacity[2] = std::exp(s[57]);
            v.scale_opacity[3] = 1.0f / (1.0f + std::exp(-s[54]));
            const gsmath::vec4 q = gsmath::normalize({s[58], s[59], s[60], s[61]});
            v.rotation[0] = q.x;
            v.rotation[1] = q.y;
            v.rotation[2] = q.z;
            v.rotation[3] = q.w;
            v.shs[0] = shs[0];
            v.shs[1] = shs[1];
            v.shs[2] = shs[2];
            // Inria stores f_rest channel-major (15 R, 15 G, 15 B); the shaders want RGB triples
            for (int j = 1; j < 16; j++) {
                v.shs[j * 3 + 0] = shs[3 + (j - 1)];
                v.shs[j * 3 + 1] = shs[3 + 15 + (j - 1)];
                v.shs[j * 3 + 2] = shs[3 + 30 + (j - 1)];
            }
        }
    };
    if (threads == 0) threads = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
    if (n < 65536 || threads == 1) {
        work(0, n);
        return;
    }
    std::vector<std::thread> pool;
    const uint64_t per = (n + threads - 1) / threads;
    for (unsigned t = 0; t < threads; t++) {
        const uint64_t b = std::min<uint64_t>(n, t * per), e = std::min<uint64_t>(n, b + per);
        if (b < e) pool.emplace_back(work, b, e);
    }
    for (auto& th : pool) th.join();
}

// Reads the body of the file with several threads (pread on disjoint slices: the page cache / NVMe queue see parallel
// requests instead of one 1-MB-at-a-time stream) and activates every slice where it was read.  The reference reads the
// whole body with one ifstream::read and activates in a single loop (GSScene.cpp:26-59).
void GSScene::loadToHost() {
    header = PlyHeader{};
    uint64_t body = 0;
    {
        std::ifstream plyFile(filename, std::ios::binary);
        loadPlyHeader(plyFile);
        body = static_cast<uint64_t>(plyFile.tellg());
    }
    const uint64_t n = getNumVertices();
    hostVertices.resize(n);
    if (n == 0) return;
    const int fd = ::open(filename.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("File does not exist: " + filename);
    const unsigned threads = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(std::thread::hardware_concurrency(), 16), (n + 65535) / 65536)));
    const uint64_t per = (n + threads - 1) / threads;
    std::vector<std::string> errors(threads);
    auto work = [&](unsigned t) {
        try {
            const uint64_t begin = std::min<uint64_t>(n, t * per), end = std::min<uint64_t>(n, begin + per);
            const uint64_t chunk = 1u << 16;  // records per pread: bounded scratch (C5-scale files are 12 GB)
            std::vector<float> records(std::min<uint64_t>(chunk, end - begin) * kRecordFloats);
            std::vector<unsigned char> raw;
            if (!layout.canonical) raw.resize(std::min<uint64_t>(chunk, end - begin) * layout.stride);
            for (uint64_t off = begin; off < end; off += chunk) {
                const uint64_t cnt = std::min(chunk, end - off);
                char* dst = layout.canonical ? reinterpret_cast<char*>(records.data()) : reinterpret_cast<char*>(raw.data());
                const uint64_t bytes = cnt * layout.stride;
                uint64_t got = 0;
                while (got < bytes) {
                    const ssize_t r = ::pread(fd, dst + got, bytes - got, static_cast<off_t>(body + off * layout.stride + got));
                    if (r <= 0) throw std::runtime_error("Unexpected end of file in " + filename);
                    got += static_cast<uint64_t>(r);
                }
                if (!layout.canonical) {  // gather the named fields into canonical 62-float records
                    for (uint64_t i = 0; i < cnt; i++) {
                        const unsigned char* src = raw.data() + i * layout.stride;
                        float* rec = records.data() + i * kRecordFloats;
                        for (int k = 0; k < 62; k++) {
                            float v = 0.0f;
                            if (layout.offset[k] >= 0) {
                                if (layout.is_double[k]) {
                                    double d;
                                    std::memcpy(&d, src + layout.offset[k], sizeof d);
                                    v = static_cast<float>(d);
                                } else {
                                    std::memcpy(&v, src + layout.offset[k], sizeof v);
                                }
                            }
                            rec[k] = v;
                        }
                    }
                }
                activateRecords(records.data(), cnt, hostVertices.data() + off, 1);
            }
        } catch (const std::exception& e) {
            errors[t] = e.what();
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    ::close(fd);
    for (const auto& e : errors)
        if (!e.empty()) throw std::runtime_error(e);
}

void GSScene::load(gsb_ctx* ctx) {
    if (!ctx) throw std::runtime_error("GSScene::load: null context");
    const auto t0 = std::chrono::steady_clock::now();
    loadToHost();
    const auto t1 = std::chrono::steady_clock::now();
    const int rc = gsb_scene_upload(ctx, reinterpret_cast<const float*>(hostVertices.data()), getNumVertices(), GSB_MEM_HOST);
    lastReadMs = std::chrono::duration<double, std::milli>(t1 - t0).count();
    lastUploadMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    if (rc != GSB_OK) throw std::runtime_error(std::string("gsb_scene_upload failed: ") + gsb_last_error(ctx));
}
