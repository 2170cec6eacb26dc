// The reference objects behind one vxo_world of oracle/_ref/libvxref.so (ref_harness.cpp).
// TEST INFRASTRUCTURE ONLY: lets C++ tests that live in the same library (the host-adapter
// self-test, tests/native/adapter_test.cpp) reach the Content / Chunks / Lighting the C API wraps.
#pragma once

#include <memory>

#include "content/Content.hpp"
#include "frontend/ContentGfxCache.hpp"
#include "graphics/render/BlocksRenderer.hpp"
#include "lighting/Lighting.hpp"
#include "settings.hpp"
#include "voxels/Chunks.hpp"
#include "world/LevelEvents.hpp"

struct MeshResult {
    bool cancelled = true;
    ChunkMeshData data;
};

struct vxo_world {
    std::unique_ptr<Content> content;
    EngineSettings settings;
    std::unique_ptr<ContentGfxCache> cache;
    LevelEvents events;
    std::unique_ptr<Chunks> chunks;
    std::unique_ptr<Lighting> lighting;
    std::unique_ptr<BlocksRenderer> renderer;
    uint64_t rendererCapacity = 0;
    MeshResult last;
};

