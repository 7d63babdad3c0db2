// ptb_bvh8.h — compressed 8-wide BVH: node layout and ray traversal (closest hit and any hit).
//
// Replaces the reference's binary BVH (`BVHNodesT`, TriangleMesh.h:6-28; traversal
// TriangleMesh.cpp:1133-1319; slab tests Geometry.h:114-204; `Triangle::intersection`
// TriangleMesh.h:82-104).  Layout after Ylitie, Karras, Laine 2017 ("Efficient Incoherent Ray
// Traversal on GPUs Through Compressed Wide BVHs"): one 80-byte node = 5 x 16-byte loads holds 8
// child boxes quantised to 8 bits on a per-node power-of-two grid.
//
//   bytes  0-11  p        float3 grid origin (node box min)
//         12-14  e        int8 exponents: cell size 2^e per axis
//            15  imask    bit s set <=> slot s holds an internal node
//         16-19  child_base   index of the first internal child (children are contiguous, slot order)
//         20-23  tri_base     index of the node's first triangle (leaf children contiguous)
//         24-26  valid24  bit 3*s + k set <=> slot s is a leaf with more than k triangles (at most 3 per leaf); the node's
//                         triangles are stored compactly in this bit order: index = tri_base + popcount(valid24 below the bit)
//            27  reserved (0)
//         28-30  hx hy hz  the same cell sizes as the exponent byte of an IEEE half: (25 + e + half_c) << 2, so that one PRMT puts a
//                          quantised plane q under it and gets the half (1024 + q) * 2^(e + half_c) (node_hitmask_h; half_c is chosen
//                          per scene, half_grid_c)
//            31  reserved (0)
//         32-79  qlo_x[8] qlo_y[8] qlo_z[8] qhi_x[8] qhi_y[8] qhi_z[8]
//
// Triangles are stored in leaf order as 3 x float4 = 48 bytes {v0, e1 = v1-v0, e2 = v2-v0}; v0.w carries
// the flags, e1.w the barycentric distance from an edge below which the reference's own arithmetic
// decides the hit (PTB_EDGE_EPS, or +inf for alpha-tested triangles; tri_exact).  Semantics mirrored from the reference: two-sided, barycentrics >= 0 inclusive,
// t >= 0 accepted, strictly nearer than the current best wins (TriangleMesh.h:82-104,
// TriangleMesh.cpp:1196-1209); alpha-mapped hits are rejected inside traversal (1198-1205).
#pragma once
#include "ptb_core.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace ptb {

struct alignas(16) Node8 {
    float px, py, pz;
    uint8_t ex, ey, ez, imask;
    uint32_t child_base, tri_base;
    uint8_t valid24[3];   // see the layout above
    uint8_t reserved0;
    uint8_t hx, hy, hz;   // half exponent bytes, see the layout above
    uint8_t reserved1;
    uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");

#if !defined(PTB_EDGE_EPS_ON)
#define PTB_EDGE_EPS_ON 1      /* 0: the fast triangle test decides everything (A/B: profiles/r02d_ab_exact_edges.txt) */
#endif
#define PTB_TRI_FLAG_ALPHA 1u  /* tri.w0 bit: this triangle's group has an alpha map that can reject */
#define PTB_TRI_FLAG_DISC 4u   /* tri.w0 bit: the triangle covers a disc of a point set (disc_cover_triangle, ptb_scene.h); tri_exact runs the disc test */
#define PTB_TRI_T_CUT (1.f - 1e-4f)   /* tri.w2 (e2.w) of an ordinary triangle: the fast test drops a hit when t * w2 >= t_best; 0 = never (volume covers, ptb_scene.h) */
#define PTB_TRI_FLAG_GHOST 2u  /* tri.w0 bit: the triangle belongs to a ghost object; shadow rays pass through it (Geometry.cpp:722) */

// ---- the half grid ---------------------------------------------------------------------------------------------------------------
// k_trace evaluates the 48 plane distances of a node with the mixed-precision FMA of sm_100 (fma.rn.f32.f16 = FHFMA: two half
// factors, float addend and result): t = half(1024 + q) * 2^(e + c)  x  half(1 / d * 2^-c)  +  float bias.  The first factor is the
// plane byte under a per-node exponent byte (hx, hy, hz of Node8), exact; the second is the ray's inverse direction rounded DOWN for
// entry planes and UP for exit planes, so that the box a ray sees is never smaller than the quantised box (q >= 0); c = half_c is one
// constant per scene that puts the cell sizes 2^e of the whole tree into the exponent range a half has next to a 11-bit integer.
// MEASURED AND NOT KEPT (profiles/r02n-r_*): 32 fewer instructions per node step (48 HADD2.F32 + 48 FFMA become 54 FHFMA), yet k_trace is
// 3 % SLOWER at equal occupancy (C2 43.1 vs 41.8 ms, C3 115.6 vs 112.5 ms per 134 M samples) and 4 % slower at the 64 registers the
// kernel then takes: FHFMA issues at the rate of one FFMA when alone (scripts/ubench/pipe_rates.cu: 0.60 vs 0.64 per cycle and
// scheduler) but not next to the min / max / select stream of the node step.  -DPTB_NODE_HALF=1 builds it; the float test is the default.
#if !defined(PTB_NODE_HALF)
#define PTB_NODE_HALF 0
#endif
#define PTB_HALF_E_HI 5       /* (1024 + 255) * 2^5 = 40928 < 65504 */
#define PTB_HALF_E_LO (-24)   /* exponent field 1: the smallest normal half under 1024 + q */
PTB_HD int half_grid_c(int e_root) { return 1 - e_root; }   // root cells (the largest) at 2^1: a re-posed scene may grow 16x before a refit has to refuse
PTB_HD uint8_t half_exp_byte(int e, int c) { return (uint8_t)((25 + e + c) << 2); }
// what the quantisation adds to a node box next to the cell-relative part: a few ulps of the coordinates
PTB_HD float node_coord_slack(float lo, float hi) { return 4e-7f * fmaxf(fabsf(lo), fabsf(hi)); }

struct Hit {
    float t, b1, b2;   // distance, barycentric of v1 (beta), of v2 (gamma)
    int32_t prim;      // triangle index in leaf order, -1 none
};

PTB_HD uint32_t highest_bit(uint32_t v) {  // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31u - (uint32_t)__clz((int)v);
#else
    return 31u - (uint32_t)__builtin_clz(v);
#endif
}
PTB_HD uint32_t popcount32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}

#if defined(__CUDA_ARCH__)
#define PTB_LDG_F4(p) __ldg(reinterpret_cast<const float4*>(p))
#else
struct float4_host { float x, y, z, w; };
#define PTB_LDG_F4(p) (*reinterpret_cast<const ptb::float4_host*>(p))
#endif

// Per-ray constants of the slab test in quantised space.
struct RayPrep {
    V3 o, d, idir;
    uint32_t oct_inv4;  // byte 0: 7 - octant (octant bit set <=> direction component negative); bytes 1..3: the slot masks of
                        // pick_slot for that octant (the half, the pairs, the single slots visited first)
};
PTB_HD RayPrep ray_prep(V3 o, V3 d) {
    RayPrep r;
    r.o = o; r.d = d;
    const float eps = 1e-30f;
    r.idir.x = 1.f / (fabsf(d.x) > eps ? d.x : copysignf(eps, d.x));
    r.idir.y = 1.f / (fabsf(d.y) > eps ? d.y : copysignf(eps, d.y));
    r.idir.z = 1.f / (fabsf(d.z) > eps ? d.z : copysignf(eps, d.z));
    const uint32_t oinv = (d.x < 0 ? 0u : 4u) | (d.y < 0 ? 0u : 2u) | (d.z < 0 ? 0u : 1u);
    // children are visited in descending (slot ^ oinv): first the half of the slots whose bit 2 differs from oinv's, ...
    r.oct_inv4 = oinv | ((oinv & 4u) ? 0x0f00u : 0xf000u) | ((oinv & 2u) ? 0x330000u : 0xcc0000u) | ((oinv & 1u) ? 0x55000000u : 0xaa000000u);
    return r;
}

// Alpha callback data: uv + group + object of a triangle, and the material table (see ptb_scene.h).
struct AlphaCtx;
PTB_HD bool alpha_rejects(const AlphaCtx* ctx, int prim, float b1, float b2);
// Triangle::intersection (TriangleMesh.h:82-104) in the reference's own arithmetic: object-space ray, plane + Gram barycentrics, every
// operation rounded separately (ptb_scene.h).  Decides the hits the fast test below cannot call: rays within PTB_EDGE_EPS of an edge.
PTB_HD bool tri_exact_available(const AlphaCtx* ctx);
#if defined(PTB_EXACT_INLINE)
#define PTB_EXACT_LINKAGE PTB_HD
#else
#define PTB_EXACT_LINKAGE PTB_HD_NOINLINE     /* out of line: the hot loop keeps its registers */
#endif
PTB_EXACT_LINKAGE bool tri_exact(const AlphaCtx* ctx, int prim, V3 o, V3 d, float tbest, float& t, float& b1, float& b2);

// Tests the 8 children of `n` against the ray; returns the hit mask in SLOT order: bit 24 + s = the internal child in slot s,
// bits 3s..3s+2 = the triangles of the leaf in slot s (only bits that exist: valid24 / imask).
#if defined(__CUDA_ARCH__)
// Device form.  ncu on the straightforward form showed the XU pipe (48 I2F.U8 per node) as the busiest pipe, so the
// quantised planes are turned into floats without any conversion instruction (see planes4 below) and the bias is folded
// into the FFMA constant.
// Both the fma pipe (FFMA, HADD2, IMAD) and the alu pipe (PRMT, FMNMX, LOP3, SHF, SEL, ISETP) issue one warp instruction
// every 2 cycles per SM sub-partition; the node test is bound by the alu pipe (about 116 alu vs 96 fma instructions per node).
#if !defined(PTB_PLANES_PRMT32)
// default: one PRMT (alu) builds a half2 {1024+q0, 1024+q1} (0x6400 | q) for TWO planes and HADD2.F32 (fma pipe) widens each;
// the bias folds into the FFMA constant: t = (1024+q)*ad + (bo - 1024*ad), at most 2^-24 * 1024 = 6e-5 of a cell of rounding.
#define PTB_PLANE_BIAS 1024.f
__device__ __forceinline__ void planes4(uint32_t w, float ad, float bo, float& t0, float& t1, float& t2, float& t3) {
    const uint32_t h01 = __byte_perm(w, 0x64646464u, 0x4140), h23 = __byte_perm(w, 0x64646464u, 0x4342);
    const __half2 a = *reinterpret_cast<const __half2*>(&h01), b = *reinterpret_cast<const __half2*>(&h23);
    t0 = fmaf(__low2float(a), ad, bo); t1 = fmaf(__high2float(a), ad, bo);
    t2 = fmaf(__low2float(b), ad, bo); t3 = fmaf(__high2float(b), ad, bo);
}
#else
// measured alternative (r01e, 1.5 % SLOWER on C2/C3): one PRMT per plane drops the byte into bits 8..15 of 0x47000000 = the
// float 32768 + q, no conversion instruction: 23 fewer instructions per node, but all 48 byte moves land on the alu pipe,
// which is the busier one.  Fold rounding 2^-24 * 32768 = 2e-3 of a cell (builder slack: 4e-3).
#define PTB_PLANE_BIAS 32768.f
__device__ __forceinline__ void planes4(uint32_t w, float ad, float bo, float& t0, float& t1, float& t2, float& t3) {
    t0 = fmaf(__uint_as_float(__byte_perm(w, 0x47000000u, 0x7404)), ad, bo);
    t1 = fmaf(__uint_as_float(__byte_perm(w, 0x47000000u, 0x7414)), ad, bo);
    t2 = fmaf(__uint_as_float(__byte_perm(w, 0x47000000u, 0x7424)), ad, bo);
    t3 = fmaf(__uint_as_float(__byte_perm(w, 0x47000000u, 0x7434)), ad, bo);
}
#endif
__device__ __forceinline__ uint32_t node_hitmask(const F4& n0, const F4& n1, const F4& n2, const F4& n3, const F4& n4, const RayPrep& r, float tmax) {
    const uint32_t e_imask = f2u(n0.w);
    const float sx = u2f(((e_imask >> 0) & 0xffu) << 23), sy = u2f(((e_imask >> 8) & 0xffu) << 23), sz = u2f(((e_imask >> 16) & 0xffu) << 23);
    const float adx = sx * r.idir.x, ady = sy * r.idir.y, adz = sz * r.idir.z;
    const float box = fmaf(-PTB_PLANE_BIAS, adx, (n0.x - r.o.x) * r.idir.x), boy = fmaf(-PTB_PLANE_BIAS, ady, (n0.y - r.o.y) * r.idir.y),
                boz = fmaf(-PTB_PLANE_BIAS, adz, (n0.z - r.o.z) * r.idir.z);
    const bool negx = r.d.x < 0, negy = r.d.y < 0, negz = r.d.z < 0;
    // The mask is built in SLOT order with immediates (slot s: its three triangle bits and its internal-child bit) and cut down to
    // what exists with one AND; the octant order is applied when a child is picked (pick_slot).  The per-child byte extracts and
    // variable shifts of a traversal-ordered mask were a quarter of the node step's instructions.
#if defined(PTB_MASK_FMA)
    uint32_t hitmask = 0xffffffffu;     // A/B: the mask accumulated on the fma pipe (FADD, IMAD.HI, IMAD per child instead of FSETP + predicated add)
#else
    uint32_t hitmask = 0;
#endif
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t lox = f2u(half ? n2.y : n2.x), loy = f2u(half ? n2.w : n2.z), loz = f2u(half ? n3.y : n3.x);
        const uint32_t hix = f2u(half ? n3.w : n3.z), hiy = f2u(half ? n4.y : n4.x), hiz = f2u(half ? n4.w : n4.z);
        float tnx[4], tny[4], tnz[4], tfx[4], tfy[4], tfz[4];
        planes4(negx ? hix : lox, adx, box, tnx[0], tnx[1], tnx[2], tnx[3]);
        planes4(negx ? lox : hix, adx, box, tfx[0], tfx[1], tfx[2], tfx[3]);
        planes4(negy ? hiy : loy, ady, boy, tny[0], tny[1], tny[2], tny[3]);
        planes4(negy ? loy : hiy, ady, boy, tfy[0], tfy[1], tfy[2], tfy[3]);
        planes4(negz ? hiz : loz, adz, boz, tnz[0], tnz[1], tnz[2], tnz[3]);
        planes4(negz ? loz : hiz, adz, boz, tfz[0], tfz[1], tfz[2], tfz[3]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float tn = fmaxf(fmaxf(tnx[j], tny[j]), fmaxf(tnz[j], 0.f));
            const float tf = fminf(fminf(tfx[j], tfy[j]), fminf(tfz[j], tmax));
            const int s = 4 * half + j;
#if defined(PTB_MASK_FMA)
            // tn <= tf  <=>  the sign of tf - tn is clear (both finite or tf = +inf; never NaN: the min / max drop NaNs and 0, tmax are numbers)
            const int32_t miss = __mulhi((int32_t)f2u(tf - tn), 2);                 // 0 or -1, IMAD.HI
            hitmask += (uint32_t)miss * ((7u << (3 * s)) | (1u << (24 + s)));        // IMAD: a missed child takes its bits out of the full mask
#else
            if (tn <= tf) hitmask |= (7u << (3 * s)) | (1u << (24 + s));
#endif
        }
    }
    return hitmask & ((e_imask & 0xff000000u) | (f2u(n1.z) & 0x00ffffffu));
}
#else
inline uint32_t node_hitmask(const F4& n0, const F4& n1, const F4& n2, const F4& n3, const F4& n4, const RayPrep& r, float tmax) {
    const uint32_t e_imask = f2u(n0.w);
    const float sx = u2f(((e_imask >> 0) & 0xffu) << 23), sy = u2f(((e_imask >> 8) & 0xffu) << 23), sz = u2f(((e_imask >> 16) & 0xffu) << 23);
    const float adx = sx * r.idir.x, ady = sy * r.idir.y, adz = sz * r.idir.z;
    const float box = (n0.x - r.o.x) * r.idir.x, boy = (n0.y - r.o.y) * r.idir.y, boz = (n0.z - r.o.z) * r.idir.z;
    const uint32_t qlox_lo = f2u(n2.x), qlox_hi = f2u(n2.y), qloy_lo = f2u(n2.z), qloy_hi = f2u(n2.w);
    const uint32_t qloz_lo = f2u(n3.x), qloz_hi = f2u(n3.y), qhix_lo = f2u(n3.z), qhix_hi = f2u(n3.w);
    const uint32_t qhiy_lo = f2u(n4.x), qhiy_hi = f2u(n4.y), qhiz_lo = f2u(n4.z), qhiz_hi = f2u(n4.w);
    const bool negx = r.d.x < 0, negy = r.d.y < 0, negz = r.d.z < 0;
    uint32_t hitmask = 0;
    for (int half = 0; half < 2; half++) {
        const uint32_t lox = half ? qlox_hi : qlox_lo, loy = half ? qloy_hi : qloy_lo, loz = half ? qloz_hi : qloz_lo;
        const uint32_t hix = half ? qhix_hi : qhix_lo, hiy = half ? qhiy_hi : qhiy_lo, hiz = half ? qhiz_hi : qhiz_lo;
        // entry plane is lo for positive direction, hi for negative
        const uint32_t nx = negx ? hix : lox, fx = negx ? lox : hix;
        const uint32_t ny = negy ? hiy : loy, fy = negy ? loy : hiy;
        const uint32_t nz = negz ? hiz : loz, fz = negz ? loz : hiz;
        for (int j = 0; j < 4; j++) {
            const uint32_t sh = 8u * j;
            const float tnx = (float)((nx >> sh) & 0xffu) * adx + box;
            const float tny = (float)((ny >> sh) & 0xffu) * ady + boy;
            const float tnz = (float)((nz >> sh) & 0xffu) * adz + boz;
            const float tfx = (float)((fx >> sh) & 0xffu) * adx + box;
            const float tfy = (float)((fy >> sh) & 0xffu) * ady + boy;
            const float tfz = (float)((fz >> sh) & 0xffu) * adz + boz;
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.f));
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));
            const int s = 4 * half + j;
            if (tn <= tf) hitmask |= (7u << (3 * s)) | (1u << (24 + s));   // slot s: its triangle bits and its internal-child bit
        }
    }
    return hitmask & ((e_imask & 0xff000000u) | (f2u(n1.z) & 0x00ffffffu));   // what exists: imask, valid24
}
#endif

#if defined(__CUDACC__)
// ---- the same test with half factors (see "the half grid" above): 48 FHFMA instead of 48 HADD2.F32 + 48 FFMA -----------------------
// d = a * b + c, a and b halves picked from half2 registers ("l_" / "h_" of the first, "m_" / "n_" of the second), c and d floats
#define PTB_FHFMA(d, a2, asel, b2, bsel, c) \
    asm("{.reg .f16 l_, h_, m_, n_; mov.b32 {l_, h_}, %1; mov.b32 {m_, n_}, %2; fma.rn.f32.f16 %0, " asel ", " bsel ", %3;}" : "=f"(d) : "r"(a2), "r"(b2), "f"(c))
// The ray's slopes on the half grid: {round-down, round-up}(1 / d * 2^-half_c) as a half2 per axis.  A component too steep for a half
// (|1 / d| * 2^-half_c > 65504: the ray runs along the planes of that axis) becomes the largest finite half on the side that has to stay
// below and an infinity on the other; the infinite side turns its plane distances into NaN (inf - inf), which the min / max of the
// interval test drop (IEEE minNum / maxNum): that side simply does not constrain the interval any more.
struct RaySlopes { uint32_t x, y, z; };
__device__ __forceinline__ uint32_t half_slope2(float idir, float scale) {
    const float v = idir * scale;
#if defined(PTB_SLOPES_CVT_DIRECTED)      // A/B: the directed conversions themselves (F2F.F16.F32.RM / .RP)
    uint16_t dn, up;
    asm("cvt.rm.f16.f32 %0, %1;" : "=h"(dn) : "f"(v));
    asm("cvt.rp.f16.f32 %0, %1;" : "=h"(up) : "f"(v));
    return (uint32_t)dn | ((uint32_t)up << 16);
#else
    // round to nearest, then step to the neighbour on the side the rounding left: halves of one sign are ordered like their bit patterns
    const __half h = __float2half_rn(v);
    const float back = __half2float(h);
    const uint32_t b = (uint32_t)__half_as_ushort(h);
    const uint32_t away = (b & 0x8000u) ? 0xffffffffu : 1u;                  // bit-pattern step that moves a half towards +inf ... (negative: b - 1)
    const uint32_t dn = (back > v) ? b - away : b;                             // ... and towards -inf.  +-0 cannot occur (|v| >= 2^-half_c), an overflow gives
    const uint32_t up = (back < v) ? b + away : b;                             // +-inf, whose inner neighbour is the largest finite half: what the directed conversion returns
    return (dn & 0xffffu) | (up << 16);
#endif
}
#define PTB_RAY_STEEP 8u     /* bit of RayPrep::oct_inv4: some component of 1 / d does not fit a half on this scene's grid (or is NaN) */
__device__ __forceinline__ RaySlopes ray_slopes(RayPrep& r, int half_c) {
    const float scale = u2f((uint32_t)(127 - half_c) << 23);
    RaySlopes s;
    s.x = half_slope2(r.idir.x, scale); s.y = half_slope2(r.idir.y, scale); s.z = half_slope2(r.idir.z, scale);
    // A ray that runs along the planes of an axis (|1 / d| * 2^-half_c beyond the largest half) would lose that axis' constraint in
    // the half form: correct, but such a ray then visits every box above and below it, and one ray that takes milliseconds holds a
    // whole persistent launch (measured, profiles/r02n_ab_node_half.txt).  These rays (a few in a million) and NaN rays, which
    // must fail every test, take the float node test.
    const float lim = 65504.f / scale;
    const float osum = r.o.x + r.o.y + r.o.z;
    if (!(fabsf(r.idir.x) <= lim && fabsf(r.idir.y) <= lim && fabsf(r.idir.z) <= lim && osum == osum)) r.oct_inv4 |= PTB_RAY_STEEP;   // (a NaN fails every comparison)
    return s;
}
// four planes of one axis: bytes of `w` under the node's half exponent byte (byte AXIS of `hw`), times the entry (FAR = false: rounded
// down) or exit slope, plus the folded bias
template <int AXIS, bool FAR>
__device__ __forceinline__ void planes4h(uint32_t w, uint32_t hw, uint32_t slope2, float bo, float& t0, float& t1, float& t2, float& t3) {
    const uint32_t h01 = __byte_perm(w, hw, 0x4040u + 0x1010u * AXIS + 0x0100u), h23 = __byte_perm(w, hw, 0x4040u + 0x1010u * AXIS + 0x0302u);
    if (FAR) {
        PTB_FHFMA(t0, h01, "l_", slope2, "n_", bo); PTB_FHFMA(t1, h01, "h_", slope2, "n_", bo);
        PTB_FHFMA(t2, h23, "l_", slope2, "n_", bo); PTB_FHFMA(t3, h23, "h_", slope2, "n_", bo);
    } else {
        PTB_FHFMA(t0, h01, "l_", slope2, "m_", bo); PTB_FHFMA(t1, h01, "h_", slope2, "m_", bo);
        PTB_FHFMA(t2, h23, "l_", slope2, "m_", bo); PTB_FHFMA(t3, h23, "h_", slope2, "m_", bo);
    }
}
// bias of an axis: t(q) = (1024 + q) * A * s + (B - 1024 * A * s) with A = 2^(e + c) (the half under the exponent byte with q = 0 is
// 1024 * A), s the half slope and B = (p - o) / d the distance of the grid origin, in float as in the float form
template <int AXIS>
__device__ __forceinline__ void bias2h(uint32_t hw, uint32_t slope2, float B, float& bo_near, float& bo_far) {
    const uint32_t a0 = __byte_perm(hw, 0u, 0x4404u + 0x0010u * AXIS);     // {1024 * A, 0}
    const uint32_t neg = slope2 ^ 0x80008000u;
    PTB_FHFMA(bo_near, a0, "l_", neg, "m_", B);
    PTB_FHFMA(bo_far, a0, "l_", neg, "n_", B);
}
__device__ __forceinline__ uint32_t node_hitmask_h(const F4& n0, const F4& n1, const F4& n2, const F4& n3, const F4& n4, const RayPrep& r, const RaySlopes& rs, float tmax) {
    const uint32_t hw = f2u(n1.w);
    float bnx, bfx, bny, bfy, bnz, bfz;
    bias2h<0>(hw, rs.x, (n0.x - r.o.x) * r.idir.x, bnx, bfx);
    bias2h<1>(hw, rs.y, (n0.y - r.o.y) * r.idir.y, bny, bfy);
    bias2h<2>(hw, rs.z, (n0.z - r.o.z) * r.idir.z, bnz, bfz);
    const bool negx = r.d.x < 0, negy = r.d.y < 0, negz = r.d.z < 0;
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const uint32_t lox = f2u(half ? n2.y : n2.x), loy = f2u(half ? n2.w : n2.z), loz = f2u(half ? n3.y : n3.x);
        const uint32_t hix = f2u(half ? n3.w : n3.z), hiy = f2u(half ? n4.y : n4.x), hiz = f2u(half ? n4.w : n4.z);
        float tnx[4], tny[4], tnz[4], tfx[4], tfy[4], tfz[4];
        planes4h<0, false>(negx ? hix : lox, hw, rs.x, bnx, tnx[0], tnx[1], tnx[2], tnx[3]);
        planes4h<0, true>(negx ? lox : hix, hw, rs.x, bfx, tfx[0], tfx[1], tfx[2], tfx[3]);
        planes4h<1, false>(negy ? hiy : loy, hw, rs.y, bny, tny[0], tny[1], tny[2], tny[3]);
        planes4h<1, true>(negy ? loy : hiy, hw, rs.y, bfy, tfy[0], tfy[1], tfy[2], tfy[3]);
        planes4h<2, false>(negz ? hiz : loz, hw, rs.z, bnz, tnz[0], tnz[1], tnz[2], tnz[3]);
        planes4h<2, true>(negz ? loz : hiz, hw, rs.z, bfz, tfz[0], tfz[1], tfz[2], tfz[3]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float tn = fmaxf(fmaxf(tnx[j], tny[j]), fmaxf(tnz[j], 0.f));
            const float tf = fminf(fminf(tfx[j], tfy[j]), fminf(tfz[j], tmax));
            const int s = 4 * half + j;
            if (tn <= tf) hitmask |= (7u << (3 * s)) | (1u << (24 + s));
        }
    }
    return hitmask & ((f2u(n0.w) & 0xff000000u) | (f2u(n1.z) & 0x00ffffffu));
}
// the node test of k_trace
#if defined(PTB_STEEP_NOINLINE)
__device__ __noinline__ uint32_t node_hitmask_steep(F4 n0, F4 n1, F4 n2, F4 n3, F4 n4, V3 o, V3 d, V3 idir, float tmax) {
    RayPrep r; r.o = o; r.d = d; r.idir = idir; r.oct_inv4 = 0;
    return node_hitmask(n0, n1, n2, n3, n4, r, tmax);
}
#endif
__device__ __forceinline__ uint32_t node_hitmask_k(const F4& n0, const F4& n1, const F4& n2, const F4& n3, const F4& n4, const RayPrep& r, const RaySlopes& rs, float tmax) {
#if defined(PTB_STEEP_NOINLINE)
    if (r.oct_inv4 & PTB_RAY_STEEP) return node_hitmask_steep(n0, n1, n2, n3, n4, r.o, r.d, r.idir, tmax);
#else
    if (r.oct_inv4 & PTB_RAY_STEEP) return node_hitmask(n0, n1, n2, n3, n4, r, tmax);
#endif
    return node_hitmask_h(n0, n1, n2, n3, n4, r, rs, tmax);
}
#endif

// Barycentric band around the triangle's edges inside which the fast test defers to tri_exact.  The fast test works on world-space
// float vertices, the reference on object-space ones: their barycentrics differ by up to a few 1e-5 (coordinates of magnitude 50
// against edges of 0.1), so a ray that close to an edge can land on either side.  A ray near a SHARED edge is near it in both
// triangles, so both get the reference's t and the reference's accept decision, and the nearer one wins as it does there.
#if !defined(PTB_EDGE_EPS)
#define PTB_EDGE_EPS 5e-4f
#endif

// Möller–Trumbore on {v0,e1,e2}; two-sided; accepts b1,b2 >= 0, b1+b2 <= 1, 0 <= t < tbest.
PTB_HD bool tri_test(const F4& a, const F4& b, const F4& c, const RayPrep& r, float tbest, float& t, float& b1, float& b2, const AlphaCtx* ex, int prim) {
    const V3 v0 = v3(a.x, a.y, a.z), e1 = v3(b.x, b.y, b.z), e2 = v3(c.x, c.y, c.z);
    const V3 pvec = cross(r.d, e2);
    const float det = dot(e1, pvec);
#if defined(__CUDA_ARCH__) && !defined(PTB_TRI_RCP_EXACT)
    // one MUFU.RCP instead of the IEEE division sequence (9 instructions + a slow path): the signs of u, v, t and hence every
    // accept test but the last ulp of `1-u-v >= 0` and `t < tbest` are unaffected; |det| below 1.2e-38 flushes to a miss.
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(det));
#else
    const float inv = 1.f / det;
#endif
    const V3 tvec = r.o - v0;
    const float u = dot(tvec, pvec) * inv;
    const V3 qvec = cross(tvec, e1);
    const float v = dot(r.d, qvec) * inv;
    const float tt = dot(e2, qvec) * inv;
    // written so that NaN (degenerate triangle, det == 0) fails every test
    const float m = fminf(fminf(u, v), 1.f - u - v);
#if PTB_EDGE_EPS_ON
    // inside the triangle grown by PTB_EDGE_EPS, in front of the origin and not clearly beyond the best hit ...
    // (c.w = PTB_TRI_T_CUT, or 0 on the covering triangles of a volume, whose own hit may be nearer than the face the ray crosses)
    if (!(m >= -PTB_EDGE_EPS) || !(tt >= 0.f) || !(tt * c.w < tbest)) return false;
    // ... and near an edge, or alpha-tested (b.w = PTB_EDGE_EPS, or +inf when the texel the uv lands on must be the reference's):
    // the reference's arithmetic decides
    if (m < b.w && tri_exact_available(ex)) return tri_exact(ex, prim, r.o, r.d, tbest, t, b1, b2);
    if (!(m >= 0.f) || !(tt < tbest)) return false;
#else
    if (!(m >= 0.f) || !(tt >= 0.f) || !(tt < tbest)) return false;
#endif
    t = tt; b1 = u; b2 = v;
    return true;
}

// The same test for the persistent traversal kernel (k_trace), which must not carry tri_exact's registers through its hot loop:
// returns 0 miss, 1 hit (t, b1, b2 set), 2 "the reference's arithmetic has to decide" (a ray within PTB_EDGE_EPS of an edge, or an
// alpha-mapped triangle, whose texel must be the reference's).  k_trace pushes case 2 on the traversal stack as a deferred
// candidate and runs tri_exact when it pops it (a few node visits later: t_best is simply not shrunk in the meantime).
PTB_HD int tri_test_classify(const F4& a, const F4& b, const F4& c, const RayPrep& r, float tbest, float& t, float& b1, float& b2) {
    const V3 v0 = v3(a.x, a.y, a.z), e1 = v3(b.x, b.y, b.z), e2 = v3(c.x, c.y, c.z);
    const V3 pvec = cross(r.d, e2);
    const float det = dot(e1, pvec);
#if defined(__CUDA_ARCH__) && !defined(PTB_TRI_RCP_EXACT)
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(det));
#else
    const float inv = 1.f / det;
#endif
    const V3 tvec = r.o - v0;
    const float u = dot(tvec, pvec) * inv;
    const V3 qvec = cross(tvec, e1);
    const float v = dot(r.d, qvec) * inv;
    const float tt = dot(e2, qvec) * inv;
    const float m = fminf(fminf(u, v), 1.f - u - v);
#if PTB_EDGE_EPS_ON
    // same instruction count as the plain accept test for the common outcome (a miss): one min3 and three compares
    if (!(m >= -PTB_EDGE_EPS) || !(tt >= 0.f) || !(tt * c.w < tbest)) return 0;   // c.w = PTB_TRI_T_CUT (t of the two formulations differs in the last bits) or 0 (yarn covers)
    if (m < b.w) return 2;            // b.w = PTB_EDGE_EPS, or +inf on alpha-tested triangles
    if (!(tt < tbest)) return 0;
#else
    if (!(m >= 0.f) || !(tt >= 0.f) || !(tt < tbest)) return 0;
#endif
    t = tt; b1 = u; b2 = v;
    return 1;
}

// The hit internal child (bits 24..31 of a node group, slot order) that the ray meets first: the slot s with the largest s ^ oinv.
// Three narrowing steps: keep the preferred half / pairs / slots whenever one of them is hit.
PTB_HD uint32_t pick_slot(uint32_t hits8, uint32_t oct_inv4) {
    uint32_t x = hits8, t;
    t = x & (oct_inv4 >> 8) & 0xffu;  x = t ? t : x;
    t = x & (oct_inv4 >> 16) & 0xffu; x = t ? t : x;
    t = x & (oct_inv4 >> 24);         x = t ? t : x;
    return highest_bit(x);            // one bit is left
}

// Traversal stack entries per ray.  A node step pushes at most two entries (the node's other hit children and the triangle group it
// postpones), so a BVH8 of `depth` wide levels needs at most 2 * depth; ptb_commit refuses deeper trees (PTB_ERR_UNSUPPORTED) instead
// of dropping subtrees the way a full stack would (the reference's own 50-entry stack overflows silently, TriangleMesh.cpp:1158).
#if !defined(PTB_STACK)
#define PTB_STACK 64
#endif

struct TraverseCounters {
    uint32_t nodes, tris;
};

// ANY_HIT: returns at the first accepted triangle with t < tmax (Scene::intersection_shadow semantics:
// the caller passes tmax = 0.999*dist_light, TriangleMesh.cpp:1239-1319 + Geometry.cpp:735).
template <bool ANY_HIT, bool COUNT>
PTB_HD bool traverse(const F4* __restrict__ nodes, const F4* __restrict__ tris, const AlphaCtx* actx, V3 o, V3 d, float tmax,
                     Hit& hit, TraverseCounters* cnt) {
    const RayPrep r = ray_prep(o, d);
    U2 stack[PTB_STACK];
    int sp = 0;
    U2 ngroup, tgroup;
    uint32_t tvalid = 0;
    // the root enters as a one-child group: bit 31 set, no imask bits -> relative index 0
    ngroup.x = 0; ngroup.y = 0x80000000u;
    bool found = false;
    float tbest = tmax;
    for (;;) {
        // invariant: ngroup has at least one pending internal child (a bit in 31..24)
        {
            const uint32_t hits_imask = ngroup.y;
            const uint32_t slot = pick_slot(hits_imask >> 24, r.oct_inv4);
            const uint32_t child_base = ngroup.x;
            ngroup.y &= ~(1u << (24u + slot));
            if (ngroup.y > 0x00ffffffu) {
                if (sp < PTB_STACK) stack[sp++] = ngroup;
            }
            const uint32_t rel = popcount32(hits_imask & ~(0xffffffffu << slot));
            const uint32_t child_index = child_base + rel;
            const F4* np = nodes + (size_t)child_index * 5;
            F4 n0, n1, n2, n3, n4;
            {
                auto l0 = PTB_LDG_F4(np + 0); auto l1 = PTB_LDG_F4(np + 1); auto l2 = PTB_LDG_F4(np + 2);
                auto l3 = PTB_LDG_F4(np + 3); auto l4 = PTB_LDG_F4(np + 4);
                n0.x = l0.x; n0.y = l0.y; n0.z = l0.z; n0.w = l0.w;
                n1.x = l1.x; n1.y = l1.y; n1.z = l1.z; n1.w = l1.w;
                n2.x = l2.x; n2.y = l2.y; n2.z = l2.z; n2.w = l2.w;
                n3.x = l3.x; n3.y = l3.y; n3.z = l3.z; n3.w = l3.w;
                n4.x = l4.x; n4.y = l4.y; n4.z = l4.z; n4.w = l4.w;
            }
            if (COUNT) cnt->nodes++;
            const uint32_t hm = node_hitmask(n0, n1, n2, n3, n4, r, tbest);
            const uint32_t imask = f2u(n0.w) >> 24;
            ngroup.x = f2u(n1.x);
            tgroup.x = f2u(n1.y);
            ngroup.y = (hm & 0xff000000u) | imask;
            tgroup.y = hm & 0x00ffffffu;
            tvalid = f2u(n1.z) & 0x00ffffffu;
        }
        while (tgroup.y != 0) {
            const uint32_t ti = highest_bit(tgroup.y);
            tgroup.y &= ~(1u << ti);
            const uint32_t prim = tgroup.x + popcount32(tvalid & ~(0xffffffffu << ti));
            const F4* tp = tris + (size_t)prim * 3;
            F4 a, b, c;
            {
                auto l0 = PTB_LDG_F4(tp + 0); auto l1 = PTB_LDG_F4(tp + 1); auto l2 = PTB_LDG_F4(tp + 2);
                a.x = l0.x; a.y = l0.y; a.z = l0.z; a.w = l0.w;
                b.x = l1.x; b.y = l1.y; b.z = l1.z; b.w = l1.w;
                c.x = l2.x; c.y = l2.y; c.z = l2.z; c.w = l2.w;
            }
            if (COUNT) cnt->tris++;
            float t, b1, b2;
            if (tri_test(a, b, c, r, tbest, t, b1, b2, actx, (int)prim)) {
                if ((f2u(a.w) & PTB_TRI_FLAG_ALPHA) && alpha_rejects(actx, (int)prim, b1, b2)) continue;
                if (ANY_HIT && (f2u(a.w) & PTB_TRI_FLAG_GHOST)) continue;
                tbest = t;
                hit.t = t; hit.b1 = b1; hit.b2 = b2; hit.prim = (int32_t)prim;
                found = true;
                if (ANY_HIT) return true;
            }
        }
        if (ngroup.y <= 0x00ffffffu) {
            if (sp > 0) ngroup = stack[--sp];
            else break;
        }
    }
    return found;
}

// Every accepted triangle with 0 <= t < tmax, in traversal order: `on_hit(prim, t, b1, b2, flags)`.  Used by the subsurface probe
// (TriMesh::reservoir_sampling_intersection, TriangleMesh.cpp:1321-1424), which is not a hot path.
template <class F>
PTB_HD void traverse_all(const F4* __restrict__ nodes, const F4* __restrict__ tris, const AlphaCtx* actx, V3 o, V3 d, float tmax, F& on_hit) {
    const RayPrep r = ray_prep(o, d);
    U2 stack[PTB_STACK];
    int sp = 0;
    U2 ngroup, tgroup;
    uint32_t tvalid = 0;
    ngroup.x = 0; ngroup.y = 0x80000000u;
    for (;;) {
        {
            const uint32_t hits_imask = ngroup.y;
            const uint32_t slot = pick_slot(hits_imask >> 24, r.oct_inv4);
            const uint32_t child_base = ngroup.x;
            ngroup.y &= ~(1u << (24u + slot));
            if (ngroup.y > 0x00ffffffu) { if (sp < PTB_STACK) stack[sp++] = ngroup; }
            const uint32_t rel = popcount32(hits_imask & ~(0xffffffffu << slot));
            const F4* np = nodes + (size_t)(child_base + rel) * 5;
            F4 n0, n1, n2, n3, n4;
            {
                auto l0 = PTB_LDG_F4(np + 0); auto l1 = PTB_LDG_F4(np + 1); auto l2 = PTB_LDG_F4(np + 2);
                auto l3 = PTB_LDG_F4(np + 3); auto l4 = PTB_LDG_F4(np + 4);
                n0.x = l0.x; n0.y = l0.y; n0.z = l0.z; n0.w = l0.w;
                n1.x = l1.x; n1.y = l1.y; n1.z = l1.z; n1.w = l1.w;
                n2.x = l2.x; n2.y = l2.y; n2.z = l2.z; n2.w = l2.w;
                n3.x = l3.x; n3.y = l3.y; n3.z = l3.z; n3.w = l3.w;
                n4.x = l4.x; n4.y = l4.y; n4.z = l4.z; n4.w = l4.w;
            }
            const uint32_t hm = node_hitmask(n0, n1, n2, n3, n4, r, tmax);
            ngroup.x = f2u(n1.x);
            tgroup.x = f2u(n1.y);
            ngroup.y = (hm & 0xff000000u) | (f2u(n0.w) >> 24);
            tgroup.y = hm & 0x00ffffffu;
            tvalid = f2u(n1.z) & 0x00ffffffu;
        }
        while (tgroup.y != 0) {
            const uint32_t ti = highest_bit(tgroup.y);
            tgroup.y &= ~(1u << ti);
            const uint32_t prim = tgroup.x + popcount32(tvalid & ~(0xffffffffu << ti));
            const F4* tp = tris + (size_t)prim * 3;
            F4 a, b, c;
            {
                auto l0 = PTB_LDG_F4(tp + 0); auto l1 = PTB_LDG_F4(tp + 1); auto l2 = PTB_LDG_F4(tp + 2);
                a.x = l0.x; a.y = l0.y; a.z = l0.z; a.w = l0.w;
                b.x = l1.x; b.y = l1.y; b.z = l1.z; b.w = l1.w;
                c.x = l2.x; c.y = l2.y; c.z = l2.z; c.w = l2.w;
            }
            float t, b1, b2;
            if (tri_test(a, b, c, r, tmax, t, b1, b2, actx, (int)prim)) {
                if ((f2u(a.w) & PTB_TRI_FLAG_ALPHA) && alpha_rejects(actx, (int)prim, b1, b2)) continue;
                on_hit((int32_t)prim, t, b1, b2);
            }
        }
        if (ngroup.y <= 0x00ffffffu) {
            if (sp > 0) ngroup = stack[--sp];
            else break;
        }
    }
}

}  // namespace ptb
