// lt_scene.cuh — device-side scene layout (HBM) and the per-CTA shared-memory copy of small scenes.
//
// Layout in HBM, per arithmetic type R (float or double), all 16-byte aligned records:
//   TriRec<R>   v0, ab=v1-v0, ac=v2-v0, n  (12 R) + material index + is_light    64 B (f32) / 112 B (f64)
//               ab/ac are formed on the host in float64 exactly as R/src/intersects.py:64-65 forms
//               them per call, so the f64 path sees the reference's bits.
//   MatRec<R>   kd[3], emission, ior, transmission, flags                           32 B / 64 B
//   LightRec<R> pos[3], area, n[3], pad, lekd[3], pad                               48 B / 96 B
//   CompatNode<R> bmin[3], bmax[3], prim_off, second, n_prims, axis  (reference LinearBVHNode)
//   CompatLeaf<R> bounds, prim_off, n_prims, kill mask — the leaves in DESCENDING node order (see
//               lt_intersect.cuh: only leaves matter once the initial descent is over)
//   LbvhNode    2 child boxes (fp32, conservatively rounded outwards) + 2 child refs   64 B = 4 x 128-bit
#pragma once
#include "lt_common.cuh"

namespace lt {

template <typename R>
struct alignas(16) TriRec {
    R v0[3], ab[3], ac[3], n[3];
    int32_t mat;
    int32_t is_light;
    int32_t pad[2];
};
template <typename R>
struct alignas(16) MatRec {
    R kd[3], emission, ior, transmission;
    int32_t flags;
    int32_t pad;
};
template <typename R>
struct alignas(16) LightRec {
    R pos[3], area, n[3], pad0, lekd[3], pad1;
};
template <typename R>
struct alignas(16) CompatNode {
    R bmin[3], bmax[3];
    int32_t prim_off, second, n_prims, axis;
};
template <typename R>
struct alignas(16) CompatLeaf {
    R bmin[3], bmax[3];
    int32_t prim_off, n_prims;
    uint64_t kill;     // list positions of the leaves this leaf marks visited through its primitive indices (Q9)
    int32_t flat_base; // position of its first primitive in the flattened (leaf-list order) primitive sequence
    int32_t pad;
};
// BVH2 node for the GPU LBVH: child boxes live in the parent so that one 64-byte fetch decides both
// children.  ref >= 0: interior node index; ref < 0: leaf holding primitive ~ref.
struct alignas(16) LbvhNode {
    float4 x;   // lo0.x lo1.x hi0.x hi1.x   one 128-bit word per axis, the two children side by side: the float32
    float4 y;   // lo0.y lo1.y hi0.y hi1.y   node step subtracts and multiplies both children at once (add/mul.f32x2)
    float4 z;   // lo0.z lo1.z hi0.z hi1.z
    int4 d;     // ref0 ref1 - -
};
#define LT_LBVH_STACK 96
static_assert(sizeof(LbvhNode) == 64, "LbvhNode must be four 128-bit words");

template <typename R>
struct SceneView {
    const TriRec<R>* tris;
    const MatRec<R>* mats;
    const LightRec<R>* lights;
    const CompatNode<R>* cnodes;
    const CompatLeaf<R>* cleaves;
    const CompatLeaf<R>* nleaves;  // NEAREST, small scenes: per-triangle bounds in primitive order (kill = 0), else null
    const LbvhNode* lbvh;         // root = node 0
    int32_t n_tris, n_mats, n_lights, n_cnodes, n_cleaves, n_lbvh_nodes;
    int32_t n_cflat;              // primitives referenced by the COMPAT leaves (flattened sequence length)
    int32_t n_nleaves;
    double bbox_min[3], bbox_max[3];   // bounds of all vertices (small scenes; +-inf when not computed): NEAREST early miss
};

}  // namespace lt
