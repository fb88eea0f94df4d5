/* examples_c_abi.c — the boundary used from plain C (no Python, no torch): build a two-sphere scene, trace
 * one ray, render a tiny image.  gcc examples_c_abi.c -Iinclude -Lcrucible_b200 -lcrucible_b200 -lm */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "crucible_gpu.h"

int main(void) {
    printf("%s, %d sm_100 device(s)\n", cr_version(), cr_device_count());
    CrScene* s = cr_scene_create(cr_device_count() > 0 ? 0 : -1);
    if (!s) { fprintf(stderr, "%s\n", cr_last_error()); return 1; }
    CrTexture tex; memset(&tex, 0, sizeof tex); tex.kind = CR_TEX_SOLID; tex.color[0] = 0.8; tex.color[1] = 0.3; tex.color[2] = 0.2;
    CrMaterial mat[2]; memset(mat, 0, sizeof mat);
    mat[0].kind = CR_MAT_LAMBERTIAN; mat[0].tex = 0; mat[0].scatter_prob = 1.0;
    mat[1].kind = CR_MAT_METAL; mat[1].albedo[0] = mat[1].albedo[1] = mat[1].albedo[2] = 0.8; mat[1].fuzz = 0.05;
    const double spheres[8] = {0, -100.5, -1, 100.0, 0, 0, -1, 0.5};
    const int32_t m[2] = {0, 1};
    if (cr_scene_add_spheres(s, spheres, m, NULL, 2) < 0 || cr_scene_set_textures(s, &tex, 1) || cr_scene_set_materials(s, mat, 2) ||
        cr_scene_commit(s)) { fprintf(stderr, "%s\n", cr_last_error()); return 1; }
    uint64_t nodes; uint32_t depth; uint64_t vis;
    cr_scene_bvh_info(s, &nodes, &depth, &vis);
    printf("BVH: %llu nodes, depth %u, %llu primitives\n", (unsigned long long)nodes, depth, (unsigned long long)vis);
    const double ray[7] = {0, 0, 0, 0, 0, -1, 0};
    CrHit hit;
    int rc = cr_trace_batch(s, ray, 1, 0.001, INFINITY, CR_PRECISION_F64, &hit);
    if (rc) { printf("no GPU path here: %s\n", cr_last_error()); cr_scene_destroy(s); return 0; }  /* no CPU fallback */
    printf("hit prim %d at t = %.17g\n", hit.prim_index, hit.t);
    CrCamera cam; memset(&cam, 0, sizeof cam);
    cam.image_width = 64; cam.image_height = 36; cam.focus_dist = 1.0; cam.frame_rate = 24; cam.shutter_angle = 180;
    cam.viewport_height = 2.0 * tan(0.5 * 1.5707963267948966); cam.viewport_width = cam.viewport_height * 64.0 / 36.0;
    cam.vup[1] = 1; cam.look_at[2] = -1; cam.samples = 16; cam.max_depth = 10;
    CrRenderOpts opts; memset(&opts, 0, sizeof opts); opts.seed = 1;
    uint8_t* img = malloc(64 * 36 * 3);
    CrStats st;
    if (cr_render(s, &cam, &opts, NULL, img, &st)) { fprintf(stderr, "%s\n", cr_last_error()); return 1; }
    printf("rendered %llu samples, %llu rays in %.2f ms; centre pixel %d %d %d\n", (unsigned long long)st.samples,
           (unsigned long long)st.rays, st.ms_total, img[(18 * 64 + 32) * 3], img[(18 * 64 + 32) * 3 + 1], img[(18 * 64 + 32) * 3 + 2]);
    free(img);
    cr_scene_destroy(s);
    return 0;
}
