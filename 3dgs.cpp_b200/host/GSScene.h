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
    // Parse + activate only (no device needed); fills `vertices()`.
    void loadToHost();

    uint64_t getNumVertices() const { return static_cast<uint64_t>(header.numVertices); }
    const PlyHeader& getHeader() const { return header; }
    const std::vector<Vertex>& vertices() const { return hostVertices; }
    void releaseHostCopy() { std::vector<Vertex>().swap(hostVertices); }

    // One 62-float PLY record -> Vertex (src/GSScene.cpp:36-59): exp(scale), sigmoid(opacity),
    // normalised quaternion, SH re-interleave.
    static void activateRecords(const float* records, uint64_t n, Vertex* out, unsigned threads = 0);

private:
    std::string filename;
    PlyHeader header;
    std::vector<Vertex> hostVertices;
    void loadPlyHeader(std::ifstream& plyFile);
};
