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
// Shared-memory copy of a float32 triangle record, padded to 80 bytes: with 64-byte records lanes that test DIFFERENT triangles
// collide on 128-bit loads (record stride 16 words: the 16-byte chunk a lane reads falls in bank group (4k + c) mod 8, two
// groups for a whole quarter-warp; ncu round 1: 23 % of the shared wavefronts were conflict replays).  Stride 20 words puts
// triangle k's chunk c in group (5k + c) mod 8 - eight distinct groups for k mod 8 distinct.
#ifndef LT_TRI_PLANES
#define LT_TRI_PLANES 1
#endif
#if LT_TRI_PLANES
// ... and since the copy is made per CTA anyway it holds the triangle in the form the float32 test wants (tri_intersect(const
// TriPad&)): the geometric unit normal `gn`, the dual edge vectors n1 = ac x N / |N|^2, n2 = N x ab / |N|^2 (N = ab x ac, so
// n1.ab = n2.ac = 1 and n1.ac = n2.ab = 0) and the parallel-ray bound `thr`; the test reads the first four 16-byte chunks,
// shading the last (`n` is the record's own normal, which a mesh may have smoothed).
struct alignas(16) TriPad {
    float v0[3], thr;
    float gn[3];
    int32_t mat;
    float n1[3];
    int32_t is_light;
    float n2[3], pad0;
    float n[3], pad1;
};
#else
struct alignas(16) TriPad : TriRec<float> {
    float pad80[4];
};
#endif
static_assert(sizeof(TriPad) == 80, "TriPad must be 80 bytes");

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
// BVH2 node of the Karras tree (LT_TRAV_LBVH2): child boxes live in the parent so that one 64-byte fetch decides both
// children.  ref >= 0: interior node index; ref < 0: leaf holding the triangle at position ~ref of the leaf-ordered arrays
// (WTri / wleaf_prims).  Boxes are EXACT float32 (rounded outwards from the float64 vertices): a 32-byte node with 16-bit planes
// on a scene-wide grid (quantum = extent / 65000 = 2.4e-4 here) was built and measured in round 2 — half the bytes, but rays that
// START on the rough mesh then enter 2.7x as many leaf boxes (tri tests 2.24 -> 6.16 per ray; the 8-wide hierarchy's per-node
// 8-bit grid has the same quantum at leaf level and the same penalty) and the twelve I2F conversions per node cost more than
// the two saved loads: 2.35 vs 3.35 Grays/s on C3.
struct alignas(16) LbvhNode {
    float4 x;   // lo0.x lo1.x hi0.x hi1.x   one 128-bit word per axis, the two children side by side: the float32
    float4 y;   // lo0.y lo1.y hi0.y hi1.y   node step subtracts and multiplies both children at once (add/mul.f32x2)
    float4 z;   // lo0.z lo1.z hi0.z hi1.z
    int4 d;     // ref0 ref1 - -
};
#define LT_LBVH_STACK 96
static_assert(sizeof(LbvhNode) == 64, "LbvhNode must be four 128-bit words");

// 8-wide compressed node of the traversal hierarchy (collapsed from the Karras tree by lt_bvh_build), 80 B = five
// 128-bit words.  Child boxes are quantised to 8 bits per plane on a per-node grid: plane = p + q * 2^e with p the node's
// grid origin and e a per-axis power of two (after Ylitie, Karras, Laine 2017).  Quantisation only ever GROWS a child box
// (lo rounded down, hi rounded up, plus a 5 % of a quantum safety margin that covers the float32 decode, see wnode_step).
//   w0: p.x p.y p.z | e_x, e_y, e_z (biased IEEE exponents), imask (bit s: slot s holds an interior child)
//   w1: child_base | tri_base | meta[0..3] | meta[4..7]
//       meta of slot s: 0 = empty; interior child: 001|11sss (bit index 24 + s of the hit word);
//       leaf: unary triangle count (001, 011, 111) << 5 | offset of its first triangle in the node's block (< 24)
//   w2: qlo_x[0..7] | qhi_x[0..7]     w3: the same for y     w4: for z
// Interior children of a node are stored contiguously from child_base in slot order; the triangles of its leaf slots are
// stored contiguously from tri_base in the leaf-ordered triangle array (WTri / wleaf_prims).
struct alignas(16) WNode {
    uint4 w0, w1, w2, w3, w4;
};
static_assert(sizeof(WNode) == 80, "WNode must be five 128-bit words");
// float32 triangle record of the wide hierarchy, in LEAF order (the triangles a node step exposes sit next to each other):
// 48 B = three 128-bit words.  area2 = |ab x ac| replaces the stored unit normal in the reference's grazing test
// (|d.n| <= 1e-7  <=>  |det| <= 1e-7 * |ab x ac|, because det = -(d . (ab x ac))); the shading normal is rebuilt from ab x ac.
struct alignas(16) WTri {
    float v0[3]; int32_t prim;
    float ab[3]; float area2;
    float ac[3]; int32_t mat_light;      // material index | is_light << 31
};
static_assert(sizeof(WTri) == 48, "WTri must be three 128-bit words");
#define LT_WBVH_LEAF_MAX 3        // triangles per leaf slot (unary count in 3 bits)
#ifndef LT_LBVH2_LEAF_MAX
#define LT_LBVH2_LEAF_MAX 2       // triangles per leaf of the binary tree (<= the slot size; lt_bvh_build flags bits 10-11 override);
                                  // C3 probe, 1/2/3: 3050 / 3198 / 3058 Mrays/s, 112 / 85 / 79 MB, 2.24 / 2.45 / 4.16 triangle tests per ray
#endif
#define LT_WBVH_MAX_DEPTH 28      // wide levels the traversal stack (shared memory, 8 B per level and lane) can hold

template <typename R>
struct SceneView {
    const TriRec<R>* tris;
    const MatRec<R>* mats;
    const LightRec<R>* lights;
    const CompatNode<R>* cnodes;
    const CompatLeaf<R>* cleaves;
    const CompatLeaf<R>* nleaves;  // NEAREST, small scenes: per-triangle bounds in primitive order (kill = 0), else null
    const LbvhNode* lbvh;         // root = node 0 (binary tree: kept for A/B runs, flag LT_BVH_KEEP_BINARY)
    const WNode* wnodes;          // 8-wide compressed hierarchy, root = node 0
    const WTri* wtris;            // float32 triangle records in leaf order
    const int32_t* wleaf_prims;   // leaf order -> primitive index (float64 traversal tests TriRec<double>[prim])
    int32_t wbvh_depth;           // wide levels (stack entries a traversal may need)
    uint32_t wmagic;              // 0x47000000, see wplane() in lt_intersect.cuh
    int32_t n_tris, n_mats, n_lights, n_cnodes, n_cleaves, n_lbvh_nodes;
    int32_t n_cflat;              // primitives referenced by the COMPAT leaves (flattened sequence length)
    int32_t n_nleaves;
    double bbox_min[3], bbox_max[3];   // bounds of all vertices (small scenes; +-inf when not computed): NEAREST early miss
};

}  // namespace lt
