// GSScene.cpp -- see GSScene.h.  Mirrors the behaviour of src/GSScene.cpp:26-68,99-149 of the
// reference: the header is scanned only for "element vertex <N>" (the property list is recorded
// but does not drive the layout), then N fixed 62-float little-endian records follow.
#include "GSScene.h"

#include <cmath>
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
            (inVertexElement ? header.vertexProperties : header.faceProperties).push_back(property);
        } else if (token == "end_header") {
            headerEnd = true;
            break;
        }
    }
    if (!headerEnd) throw std::runtime_error("Could not find end of header");
    if (header.numVertices < 0) throw std::runtime_error("Negative vertex count in " + filename);
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

void GSScene::loadToHost() {
    header = PlyHeader{};
    std::ifstream plyFile(filename, std::ios::binary);
    loadPlyHeader(plyFile);
    const uint64_t n = getNumVertices();
    hostVertices.resize(n);
    // stream the records in bounded chunks: C5-scale files are 12 GB
    const uint64_t chunk = 1u << 20;
    std::vector<float> records(std::min(n, chunk) * kRecordFloats);
    for (uint64_t off = 0; off < n; off += chunk) {
        const uint64_t cnt = std::min(chunk, n - off);
        plyFile.read(reinterpret_cast<char*>(records.data()), static_cast<std::streamsize>(cnt * kRecordFloats * sizeof(float)));
        if (static_cast<uint64_t>(plyFile.gcount()) != cnt * kRecordFloats * sizeof(float))
            throw std::runtime_error("Unexpected end of file in " + filename);
        activateRecords(records.data(), cnt, hostVertices.data() + off);
    }
}

void GSScene::load(gsb_ctx* ctx) {
    if (!ctx) throw std::runtime_error("GSScene::load: null context");
    loadToHost();
    const int rc = gsb_scene_upload(ctx, reinterpret_cast<const float*>(hostVertices.data()), getNumVertices(), GSB_MEM_HOST);
    if (rc != GSB_OK) throw std::runtime_error(std::string("gsb_scene_upload failed: ") + gsb_last_error(ctx));
}
