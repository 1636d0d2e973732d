// GSScene -- host-side scene loader with the reference's surface (src/GSScene.h:23-53):
// construct from a path (throws if missing), load() reads the Inria-format .ply, applies the
// activations and uploads.  The Vulkan upload + precomp_cov3d dispatch (GSScene.cpp:61,157-184)
// is replaced by one gsb_scene_upload() call into libgsb200.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "gs_b200.h"

struct PlyProperty {
    std::string type;
    std::string name;
};

// Where the 62 canonical floats of one record (x y z | nx ny nz | f_dc 0-2 | f_rest 0-44 | opacity | scale 0-2 |
// rot 0-3, src/GSScene.cpp:17-24) sit inside one on-disk vertex record of the file at hand.
struct PlyLayout {
    bool canonical = true;      // the file IS 62 packed floats in the canonical order: read records verbatim
    uint64_t stride = 62 * 4;   // bytes per on-disk vertex record
    int32_t offset[62];         // byte offset of canonical float k inside a record, -1 = absent (reads as 0)
    uint8_t is_double[62];      // the source property is a 64-bit float
    int shDegree = 3;           // highest SH degree present in the file (f_rest count 0 / 9 / 24 / 45)
};

struct PlyHeader {
    std::string format;
    int64_t numVertices = 0;
    int64_t numFaces = 0;
    std::vector<PlyProperty> vertexProperties;
    std::vector<PlyProperty> faceProperties;
};

class GSScene {
public:
    // GSScene::Vertex, src/GSScene.h:41-46 (60 floats, what gsb_scene_upload consumes)
    struct Vertex {
        float position[4];
        float scale_opacity[4];
        float rotation[4];
        float shs[48];
    };
    static_assert(sizeof(Vertex) == 60 * sizeof(float), "Vertex layout");

    explicit GSScene(const std::string& filename);  // throws std::runtime_error if the file is missing

    // Parse + activate + upload to `ctx` (replaces load(const std::shared_ptr<VulkanContext>&)).
    void load(gsb_ctx* ctx);
    // wall-clock of the last load(): file read + activation (loadToHost), and gsb_scene_upload (pinned ring + cov3D ingest)
    double lastReadMs = 0.0, lastUploadMs = 0.0;
    // Parse + activate only (no device needed); fills `vertices()`.
    void loadToHost();

    uint64_t getNumVertices() const { return static_cast<uint64_t>(header.numVertices); }
    const PlyHeader& getHeader() const { return header; }
    // Layout derived from the header's property list.  The reference ignores the list and assumes
    // the canonical record (src/GSScene.cpp:119-123,36-59); files with that list -- or with none -- take the same
    // verbatim path here.  Other lists (extra properties, another order, SH degree < 3, double fields) are gathered
    // by name; a missing required field, a list property or a non-little-endian-binary format is an error.
    const PlyLayout& getLayout() const { return layout; }
    const std::vector<Vertex>& vertices() const { return hostVertices; }
    void releaseHostCopy() { std::vector<Vertex>().swap(hostVertices); }

    // One 62-float PLY record -> Vertex (src/GSScene.cpp:36-59): exp(scale), sigmoid(opacity),
    // normalised quaternion, SH re-interleave.
    static void activateRecords(const float* records, uint64_t n, Vertex* out, unsigned threads = 0);

private:
    std::string filename;
    PlyHeader header;
    PlyLayout layout;
    std::vector<Vertex> hostVertices;
    void resolveLayout();
    void loadPlyHeader(std::ifstream& plyFile);
};
