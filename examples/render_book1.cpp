// examples/render_book1.cpp — the reference's `crucible --file out --world 1` (src/main.rs:35-79) on the GPU
// backend, through the C++ host mirror.   g++ -std=c++17 examples/render_book1.cpp -Iinclude -Lcrucible_b200 -lcrucible_b200
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "crucible.hpp"

int main(int argc, char** argv) {
    std::string file = "out";
    uint32_t width = 400, samples = 100;
    uint64_t seed = 1;
    bool describe_only = false;
    for (int i = 1; i < argc; ++i) {
        auto next = [&](const char* flag) { if (i + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", flag); std::exit(2); } return argv[++i]; };
        if (!std::strcmp(argv[i], "--file")) file = next("--file");
        else if (!std::strcmp(argv[i], "--width")) width = (uint32_t)std::atoi(next("--width"));
        else if (!std::strcmp(argv[i], "--samples")) samples = (uint32_t)std::atoi(next("--samples"));
        else if (!std::strcmp(argv[i], "--seed")) seed = (uint64_t)std::atoll(next("--seed"));
        else if (!std::strcmp(argv[i], "--describe-only")) describe_only = true;
    }
    try {
        crucible::Scene scene = crucible::demo_images::book1_end_scene(0, seed, width, samples);
        if (describe_only) {  // host-only: build the reference BVH and report it (works without a GPU)
            CrScene* s = scene.flatten(-1);
            uint64_t nodes, vis;
            uint32_t depth;
            cr_scene_bvh_info(s, &nodes, &depth, &vis);
            std::printf("prims %zu nodes %llu depth %u visible %llu\n", scene.n_prims(), (unsigned long long)nodes, depth, (unsigned long long)vis);
            cr_scene_destroy(s);
            return 0;
        }
        const CrStats st = scene.render_scene(file);
        std::printf("%llu samples, %llu rays, %.2f ms on the device: %.1f Msamples/s\n", (unsigned long long)st.samples,
                    (unsigned long long)st.rays, st.ms_total, st.samples / st.ms_total / 1e3);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
