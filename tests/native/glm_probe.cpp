// Prints the bit patterns of the GLM-shim arithmetic the reference's mesher reaches
// (glm::normalize / dot / cross at BlocksRenderer.cpp:131-136,169,299-313).
// Built and run by tests/test_glm_shim.py against oracle/ref_build/glm/glm.hpp with the
// reference build's flags (-O3, no FMA contraction).  Test infrastructure only.
#include <cstdint>
#include <cstdio>
#include <cstring>

#include <glm/glm.hpp>

static uint32_t bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

int main() {
    const glm::vec3 SUN(0.2275f, 0.9388f, -0.1005f);  // BlocksRenderer.cpp:12
    const glm::vec3 axes[6] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (const auto& a : axes) {
        float d = glm::dot(glm::normalize(a), SUN);  // :131-132
        d = 0.7f + d * 0.3f;
        std::printf("sun %08x\n", bits(d));
    }
    const float scales[6] = {0.2f, 0.125f, 0.5f, 1.0f, 2.0f, 0.85f};
    for (float s : scales) {
        volatile float vs = s;  // keep the arithmetic at run time
        const glm::vec3 n = glm::normalize(glm::vec3(0, vs, 0));
        std::printf("norm %08x\n", bits(n.y));
    }
    {
        volatile float a = 0.3f, b = 0.7f, c = -0.2f;
        const glm::vec3 x(a, b, c), y(c, a, b);
        const glm::vec3 cr = glm::cross(x, y);
        std::printf("cross %08x %08x %08x\n", bits(cr.x), bits(cr.y), bits(cr.z));
        std::printf("dot %08x\n", bits(glm::dot(x, y)));
        const glm::vec3 n = glm::normalize(x);
        std::printf("normv %08x %08x %08x\n", bits(n.x), bits(n.y), bits(n.z));
    }
    {
        const glm::ivec3 t = glm::vec3(1.9f, -1.9f, 0.99999994f);  // implicit truncating conversion
        std::printf("trunc %d %d %d\n", t.x, t.y, t.z);
    }
    return 0;
}
