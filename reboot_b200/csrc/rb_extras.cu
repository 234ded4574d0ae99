// rb_extras.cu -- the callers either side of the physics step:
//   * frustum culling on the world's spatial structure: GeometryMath::getFrustumPlanes / frustumAABBDetection
//     (math/src/GeometryMath.cpp:630-757) and OSP::getVisibleFrustumCulling (physics/src/OSP.cpp:125-180)
//   * collider ingest: FbxLoader::_buildGeometryData (model/src/FbxLoader.cpp:1109-1216) from a neutral
//     indexed-mesh input (the FBX SDK itself is not available)
// Same arithmetic contract as the step kernels: -fmad=false on the device, -ffp-contract=off on the host,
// expression order of the reference.
#include "rb_internal.h"
#define RB_BLOCK 256
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

// ---- device: one thread per box. mode 0: boxes are cubes (centre, length, height, width) as in OSP.cpp:147-156;
// mode 1: boxes are the non-empty columns of the static triangle grid (x/z extent of the column, y range of its
// triangles). Visible boxes are appended (unordered) as (|centre|, index); the host sorts them front to back. ----
struct RbCullOut { float key; uint32_t idx; };

__device__ __forceinline__ bool rb_box_in_frustum(const float* __restrict__ pl, f3 mn, f3 mx) {
    // GeometryMath::frustumAABBDetection: culled iff all 8 corners lie on the positive side of one plane
    const f3 c[8] = { mk3(mn.x, mn.y, mn.z), mk3(mn.x, mn.y, mx.z), mk3(mn.x, mx.y, mx.z), mk3(mx.x, mx.y, mx.z),
                      mk3(mx.x, mx.y, mn.z), mk3(mx.x, mn.y, mn.z), mk3(mx.x, mn.y, mx.z), mk3(mn.x, mx.y, mn.z) };
    bool outside = false;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const f3 n = mk3(pl[4 * p], pl[4 * p + 1], pl[4 * p + 2]);
        const float dist = pl[4 * p + 3];
        const f3 on = mul3(n, dist);                       // "pointOnPlane = plane * distance"
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const f3 v = sub3(c[k], on);
            if ((n.x * v.x + n.y * v.y) + n.z * v.z > 0.0f) cnt++;
        }
        if (cnt == 8) outside = true;
    }
    return !outside;
}

struct RbCullPlanes { float p[24]; };

__global__ void __launch_bounds__(RB_BLOCK) rb_frustum_cull_kernel(RbCullPlanes planes, int mode, uint32_t n, const float* __restrict__ cubes6,
                                                                  RbParams P, RbCullOut* __restrict__ out, uint32_t* __restrict__ counter, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool vis = false;
    float key = 0.f;
    if (i < n) {
        f3 mn, mx, ctr;
        bool live = true;
        if (mode == 0) {
            const float* q = cubes6 + 6ull * i;
            ctr = mk3(q[0], q[1], q[2]);
            mn = mk3(q[0] - q[3] / 2.0f, q[1] - q[4] / 2.0f, q[2] - q[5] / 2.0f);      // OSP.cpp:150-155
            mx = mk3(q[0] + q[3] / 2.0f, q[1] + q[4] / 2.0f, q[2] + q[5] / 2.0f);
        } else {
            live = P.tcol_start[i + 1] > P.tcol_start[i];                            // "if (octNode->getTriangles()->size() != 0)"
            const uint32_t cx = i % P.tcx, cz = i / P.tcx;
            const float wcol = (2.0f * P.half) / (float)P.tcx;
            const float2 yr = live ? P.tcol_yr[i] : make_float2(0.f, 0.f);
            mn = mk3(-P.half + (float)cx * wcol, yr.x, -P.half + (float)cz * wcol);
            mx = mk3(-P.half + (float)(cx + 1u) * wcol, yr.y, -P.half + (float)(cz + 1u) * wcol);
            ctr = mk3((mn.x + mx.x) / 2.0f, (mn.y + mx.y) / 2.0f, (mn.z + mx.z) / 2.0f);
        }
        if (live && rb_box_in_frustum(planes.p, mn, mx)) { vis = true; key = sqrtf((ctr.x * ctr.x + ctr.y * ctr.y) + ctr.z * ctr.z); }   // center.getMagnitude(), OSP.cpp:160
    }
    // warp-aggregated append
    const uint32_t m = __ballot_sync(0xffffffffu, vis);
    if (m) {
        const uint32_t lane = threadIdx.x & 31;
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(counter, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (vis) {
            const uint32_t slot = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (slot < cap) { out[slot].key = key; out[slot].idx = i; }
        }
    }
}

// cells of the static triangle grid as boxes (min xyz, max xyz) + triangle entries per cell
__global__ void __launch_bounds__(RB_BLOCK) rb_static_cells_kernel(RbParams P, float* __restrict__ boxes6, uint32_t* __restrict__ ntri) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.tcx * P.tcz) return;
    const uint32_t cnt = P.tcol_start[i + 1] - P.tcol_start[i];
    const uint32_t cx = i % P.tcx, cz = i / P.tcx;
    const float wcol = (2.0f * P.half) / (float)P.tcx;
    const float2 yr = cnt ? P.tcol_yr[i] : make_float2(0.f, 0.f);
    float* b = boxes6 + 6ull * i;
    b[0] = -P.half + (float)cx * wcol; b[1] = yr.x; b[2] = -P.half + (float)cz * wcol;
    b[3] = -P.half + (float)(cx + 1u) * wcol; b[4] = yr.y; b[5] = -P.half + (float)(cz + 1u) * wcol;
    ntri[i] = cnt;
}

// ---- host arithmetic, expression for expression as the reference -----------------------------------------
namespace {
struct M4 { float m[16]; };
M4 m_identity() { M4 r; memset(&r, 0, sizeof r); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
M4 m_mul(const M4& a, const M4& b) {            // Matrix::operator*(Matrix), math/src/Matrix.cpp:136-166
    M4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            volatile float s = a.m[4 * i] * b.m[j] + a.m[4 * i + 1] * b.m[4 + j];
            s = s + a.m[4 * i + 2] * b.m[8 + j];
            s = s + a.m[4 * i + 3] * b.m[12 + j];
            r.m[4 * i + j] = s;
        }
    return r;
}
void m_mulv(const M4& a, const float* v, float w, float out[4]) {    // Matrix::operator*(Vector4), Matrix.cpp:124-134
    for (int i = 0; i < 4; i++) {
        volatile float s = a.m[4 * i] * v[0] + a.m[4 * i + 1] * v[1];
        s = s + a.m[4 * i + 2] * v[2];
        s = s + a.m[4 * i + 3] * w;
        out[i] = s;
    }
}
// translation * (rotX * rotY * rotZ) * scale, degrees (FbxLoader.cpp:1114-1131, Matrix.cpp:261-376)
M4 trs_matrix(const float* trs9) {
    const float k = 3.14159265f / 180.0f;       // Matrix.h:27-28
    const float tx = trs9[3] * k, ty = trs9[4] * k, tz = trs9[5] * k;
    M4 rx = m_identity(), ry = m_identity(), rz = m_identity(), sc = m_identity(), t = m_identity();
    rx.m[5] = cosf(tx); rx.m[6] = sinf(tx); rx.m[9] = -sinf(tx); rx.m[10] = cosf(tx);
    ry.m[0] = cosf(ty); ry.m[2] = -sinf(ty); ry.m[8] = sinf(ty); ry.m[10] = cosf(ty);
    rz.m[0] = cosf(tz); rz.m[1] = sinf(tz); rz.m[4] = -sinf(tz); rz.m[5] = cosf(tz);
    sc.m[0] = trs9[6]; sc.m[5] = trs9[7]; sc.m[10] = trs9[8];
    t.m[3] = trs9[0]; t.m[7] = trs9[1]; t.m[11] = trs9[2];
    return m_mul(m_mul(t, m_mul(m_mul(rx, ry), rz)), sc);
}
bool idx_ok(uint32_t nverts, uint32_t nidx, const uint32_t* idx) { for (uint32_t i = 0; i < nidx; i++) if (idx[i] >= nverts) return false; return true; }
}

extern "C" {

// GeometryMath::getFrustumPlanes (math/src/GeometryMath.cpp:630-731): left, right, far, near, bottom, top as (n, d), all
// four components divided by |n|
int rb_frustum_planes(const float inv_view_proj[16], float planes24[24]) {
    if (!inv_view_proj || !planes24) return RB_ERR_INVALID;
    static const float ndc[8][3] = { { -1, -1, 1 }, { -1, 1, 1 }, { -1, -1, -1 }, { -1, 1, -1 }, { 1, -1, 1 }, { 1, 1, 1 }, { 1, -1, -1 }, { 1, 1, -1 } };
    static const int pl[6][3] = { { 0, 2, 1 }, { 4, 5, 6 }, { 6, 3, 2 }, { 4, 0, 1 }, { 0, 4, 2 }, { 5, 1, 3 } };   // origin, A's point, B's point
    M4 m; memcpy(m.m, inv_view_proj, sizeof m.m);
    float pt[8][3];
    for (int i = 0; i < 8; i++) {
        float q[4]; m_mulv(m, ndc[i], 1.0f, q);
        pt[i][0] = q[0] / q[3]; pt[i][1] = q[1] / q[3]; pt[i][2] = q[2] / q[3];
    }
    for (int p = 0; p < 6; p++) {
        const float* o = pt[pl[p][0]];
        volatile float ax = pt[pl[p][1]][0] - o[0], ay = pt[pl[p][1]][1] - o[1], az = pt[pl[p][1]][2] - o[2];
        volatile float bx = pt[pl[p][2]][0] - o[0], by = pt[pl[p][2]][1] - o[1], bz = pt[pl[p][2]][2] - o[2];
        // C = B x A (Vector4::crossProduct, Vector4.cpp:106-117)
        volatile float cx = (by * az) - (bz * ay);
        volatile float cy = -(bx * az) + (bz * ax);
        volatile float cz = (bx * ay) - (by * ax);
        volatile float s = cx * -o[0] + cy * -o[1];
        s = s + cz * -o[2];
        volatile float d = -s;
        volatile float q = (cx * cx) + (cy * cy);
        q = q + (cz * cz);
        const float mag = sqrtf(q);
        planes24[4 * p] = cx / mag; planes24[4 * p + 1] = cy / mag; planes24[4 * p + 2] = cz / mag; planes24[4 * p + 3] = d / mag;
    }
    return RB_OK;
}

static int cull_common(rb_world* w, const float planes24[24], int mode, uint32_t n, const float* cubes6_host,
                       uint32_t* visible, uint32_t cap, uint32_t* n_visible) {
    if (!w || !planes24 || !n_visible || (cap && !visible)) return RB_ERR_INVALID;
    CU(cudaSetDevice(w->cfg.device));
    *n_visible = 0;
    if (n == 0) return RB_OK;
    // scratch lives with the world and only grows: culling is a per-frame call of a renderer
    cudaError_t e = cudaSuccess;
    if (w->cull_cap < n) {
        cudaFree(w->d_cull_out); cudaFree(w->d_cull_cubes); w->d_cull_out = nullptr; w->d_cull_cubes = nullptr; w->cull_cap = 0;
        e = cudaMalloc(&w->d_cull_out, (size_t)n * sizeof(RbCullOut));
        if (e == cudaSuccess) e = cudaMalloc(&w->d_cull_cubes, (size_t)n * 6 * sizeof(float));
        if (e == cudaSuccess && !w->d_cull_cnt) e = cudaMalloc(&w->d_cull_cnt, sizeof(uint32_t));
        if (e == cudaSuccess) w->cull_cap = n;
    }
    float* d_cubes = (float*)w->d_cull_cubes; RbCullOut* d_out = (RbCullOut*)w->d_cull_out; uint32_t* d_cnt = w->d_cull_cnt;
    if (e == cudaSuccess && mode == 0) e = cudaMemcpyAsync(d_cubes, cubes6_host, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, w->stream);
    if (e == cudaSuccess) { rb_launch_clear(nullptr, 0, d_cnt, 1, w->stream); w->launches++; }
    std::vector<RbCullOut> h;
    uint32_t cnt = 0;
    if (e == cudaSuccess) {
        RbCullPlanes pl; memcpy(pl.p, planes24, sizeof pl.p);
        rb_frustum_cull_kernel<<<(n + RB_BLOCK - 1) / RB_BLOCK, RB_BLOCK, 0, w->stream>>>(pl, mode, n, d_cubes, w->P, d_out, d_cnt, n);
        w->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, w->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
        if (e == cudaSuccess && cnt) { h.resize(cnt); e = cudaMemcpy(h.data(), d_out, (size_t)cnt * sizeof(RbCullOut), cudaMemcpyDeviceToHost); }
    }
    if (e != cudaSuccess) return fail(w, RB_ERR_CUDA, "frustum culling: %s", cudaGetErrorString(e));
    // front to back by |centre| (OSP.cpp:166-178); equal keys in index order (the reference's std::sort leaves ties unspecified)
    std::sort(h.begin(), h.end(), [](const RbCullOut& a, const RbCullOut& b) { return a.key < b.key || (a.key == b.key && a.idx < b.idx); });
    *n_visible = cnt;
    for (uint32_t i = 0; i < cnt && i < cap; i++) visible[i] = h[i].idx;
    if (cnt > cap && cap) return fail(w, RB_ERR_CAPACITY, "frustum culling: %u visible boxes, room for %u", cnt, cap);
    return RB_OK;
}

// OSP::getVisibleFrustumCulling for caller-supplied boxes given as the reference's Cube (centre xyz, length, height,
// width): 0-based indices of the boxes that meet the frustum, front to back
int rb_frustum_cull_cubes(rb_world* w, const float planes24[24], uint32_t n, const float* cubes6, uint32_t* visible, uint32_t cap, uint32_t* n_visible) {
    if (n && !cubes6) return RB_ERR_INVALID;
    return cull_common(w, planes24, 0, n, cubes6, visible, cap, n_visible);
}
// the same over the world's own static structure: the non-empty cells of the triangle column grid (cell ids, see rb_get_static_cells)
int rb_frustum_cull_static(rb_world* w, const float inv_view_proj[16], uint32_t* cells, uint32_t cap, uint32_t* n_visible) {
    if (!w || !w->built) return w ? fail(w, RB_ERR_INVALID, "rb_frustum_cull_static before rb_build") : RB_ERR_INVALID;
    float planes[24];
    int r = rb_frustum_planes(inv_view_proj, planes); if (r) return r;
    return cull_common(w, planes, 1, w->P.tcx * w->P.tcz, nullptr, cells, cap, n_visible);
}
// cells of the static structure: boxes6 = (min xyz, max xyz) per cell (tri_cols_x * tri_cols_z of them, row-major in z),
// ntri = triangles registered in the cell (0: empty, box y range 0)
int rb_get_static_cells(rb_world* w, float* boxes6, uint32_t* ntri) {
    if (!w || !w->built || !boxes6 || !ntri) return RB_ERR_INVALID;
    CU(cudaSetDevice(w->cfg.device));
    const uint32_t n = w->P.tcx * w->P.tcz;
    if (!n) return RB_OK;
    if (!w->d_cells_box) { ALLOC(w->d_cells_box, (size_t)n * 6); ALLOC(w->d_cells_n, n); }      // (the static structure never changes: kept with the world)
    float* d_b = w->d_cells_box; uint32_t* d_n = w->d_cells_n;
    rb_static_cells_kernel<<<(n + RB_BLOCK - 1) / RB_BLOCK, RB_BLOCK, 0, w->stream>>>(w->P, d_b, d_n); w->launches++;
    cudaError_t e = cudaMemcpyAsync(boxes6, d_b, (size_t)n * 6 * sizeof(float), cudaMemcpyDeviceToHost, w->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ntri, d_n, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, w->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
    if (e != cudaSuccess) return fail(w, RB_ERR_CUDA, "rb_get_static_cells: %s", cudaGetErrorString(e));
    return RB_OK;
}

// ---- static triangle column grid, built on the device (replaces the per-leaf triangle sets of OSP::generateGeometryOSP,
// physics/src/OSP.cpp:25-60, 242-332; built once: triangles never move) -------------------------------------------------
// A triangle is registered in the columns of its box's column range (half-open at the upper boundary with a 2.5e-4
// tolerance, < the smallest candidate margin of 1e-3, so an aligned lattice registers one column per cell); when that range
// spans several columns, only in those its surface actually crosses: the reference's own triangle / box predicate
// (GeometryMath::triangleCubeDetection, the test OSP uses to hand triangles down its tree) against the column prism grown by
// a slack far above the predicate's rounding. A sphere that touches the triangle touches it at a point of some column its
// box overlaps, and the triangle crosses that column: nothing is lost.
struct RbTriGrid { uint32_t n; float half, inv_w, cw, slack; };
__device__ __forceinline__ void tgrid_range(const RbTriGrid& G, float lo, float hi, uint32_t& c0, uint32_t& c1) {
    c0 = rb_col(lo, G.half, G.inv_w, G.n);
    c1 = rb_col(hi, G.half, G.inv_w, G.n);
    const uint32_t c1t = rb_col(hi - 2.5e-4f, G.half, G.inv_w, G.n);
    if (c1 > c0 && c1t < c1) c1 = c1t;
}
__device__ __forceinline__ bool tgrid_crosses(const RbTriGrid& G, f3 a, f3 b, f3 c, uint32_t x, uint32_t z, bool multi) {
    if (!multi) return true;
    const f3 ctr = mk3(-G.half + ((float)x + 0.5f) * G.cw, 0.0f, -G.half + ((float)z + 0.5f) * G.cw);
    // Cube(length, height, width): height pairs with x, width with y, length with z (GeometryMath.cpp:238-240)
    return triangle_cube_detect(a, b, c, ctr, G.cw + 2.0f * G.slack, G.cw + 2.0f * G.slack, 16.0f * G.half);
}
// pass 1: device triangle records (vertices + the static normal of GeometryMath::sphereTriangleResolution, :499-503:
// normalize((C - A) x (B - A)) in the reference's rounding order, w lanes) and the per-column counts
// pass 2 (fill != NULL): the entries -- triangle id + its box -- at start[col] + arrival rank
__global__ void __launch_bounds__(RB_BLOCK) rb_tgrid_scatter_kernel(RbTriGrid G, const float* __restrict__ xyz9, uint32_t nt, float4* __restrict__ tri,
                                                                   uint32_t* __restrict__ count, const uint32_t* __restrict__ start,
                                                                   uint32_t* __restrict__ list, float4* __restrict__ ent) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const float* p = xyz9 + 9ull * t;
    const f3 a = mk3(p[0], p[1], p[2]), b = mk3(p[3], p[4], p[5]), c = mk3(p[6], p[7], p[8]);
    if (tri) {
        const f3 nrm = normalize3(cross3(sub3(c, a), sub3(b, a)));
        tri[3ull * t] = make_float4(a.x, a.y, a.z, nrm.x); tri[3ull * t + 1] = make_float4(b.x, b.y, b.z, nrm.y); tri[3ull * t + 2] = make_float4(c.x, c.y, c.z, nrm.z);
    }
    const float bx0 = fminf(a.x, fminf(b.x, c.x)), bx1 = fmaxf(a.x, fmaxf(b.x, c.x));
    const float by0 = fminf(a.y, fminf(b.y, c.y)), by1 = fmaxf(a.y, fmaxf(b.y, c.y));
    const float bz0 = fminf(a.z, fminf(b.z, c.z)), bz1 = fmaxf(a.z, fmaxf(b.z, c.z));
    uint32_t x0, x1, z0, z1; tgrid_range(G, bx0, bx1, x0, x1); tgrid_range(G, bz0, bz1, z0, z1);
    const bool multi = x1 > x0 || z1 > z0;
    for (uint32_t z = z0; z <= z1; ++z)
        for (uint32_t x = x0; x <= x1; ++x) {
            if (!tgrid_crosses(G, a, b, c, x, z, multi)) continue;
            const size_t col = (size_t)z * G.n + x;
            const uint32_t r = atomicAdd(&count[col], 1u);
            if (list) {
                const uint32_t k = start[col] + r;
                list[k] = t;
                ent[2ull * k] = make_float4(bx0, bx1, bz0, bz1);
                ent[2ull * k + 1] = make_float4(by0, by1, __uint_as_float(t), 0.f);
            }
        }
}
// pass 3: one thread per column -- its entries into ascending triangle id (the canonical order of the narrowphase walks) and
// the y range of what it holds
__global__ void __launch_bounds__(RB_BLOCK) rb_tgrid_finish_kernel(uint32_t ncol, const uint32_t* __restrict__ start, uint32_t* __restrict__ list,
                                                                  float4* __restrict__ ent, float2* __restrict__ yr) {
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const uint32_t s = start[col], e = start[col + 1];
    float lo = INFINITY, hi = -INFINITY;
    for (uint32_t i = s + 1; i < e; ++i) {      // insertion sort: columns hold a handful of entries, arrival order is nearly sorted
        const uint32_t t = list[i]; const float4 e0 = ent[2ull * i], e1 = ent[2ull * i + 1];
        uint32_t j = i;
        while (j > s && list[j - 1] > t) { list[j] = list[j - 1]; ent[2ull * j] = ent[2ull * (j - 1)]; ent[2ull * j + 1] = ent[2ull * (j - 1) + 1]; --j; }
        list[j] = t; ent[2ull * j] = e0; ent[2ull * j + 1] = e1;
    }
    for (uint32_t i = s; i < e; ++i) { const float4 e1 = ent[2ull * i + 1]; lo = fminf(lo, e1.x); hi = fmaxf(hi, e1.y); }
    yr[col] = make_float2(lo, hi);
}
// host driver (called by rb_build): n columns per axis, triangles in w->h_tri
int rb_build_triangle_grid_device(rb_world* w, uint32_t n) {
    RbParams& P = w->P;
    const uint32_t nt = P.nt;
    const float half = w->cfg.world_half;
    const size_t ncol = (size_t)n * n;
    RbTriGrid G; G.n = n; G.half = half; G.inv_w = P.t_inv_w; G.cw = 2.0f * half / (float)n; G.slack = std::max(1e-3f * G.cw, 5e-4f);
    float* d_xyz = nullptr; uint32_t *d_count = nullptr, *d_tiles = nullptr;
    float4* d_tri; uint32_t* d_start; float2* d_yr;
    ALLOC(d_tri, (size_t)std::max<uint32_t>(nt, 1) * 3); ALLOC(d_start, ncol + 1); ALLOC(d_yr, ncol);
    CU(cudaMalloc(&d_xyz, std::max<size_t>((size_t)nt * 9, 1) * sizeof(float)));
    CU(cudaMalloc(&d_count, (ncol + 1) * sizeof(uint32_t)));
    CU(cudaMalloc(&d_tiles, (rb_scan_tiles((uint32_t)ncol) + 1) * sizeof(uint32_t)));
    auto cleanup = [&] { cudaFree(d_xyz); cudaFree(d_count); cudaFree(d_tiles); };
    cudaError_t e = cudaMemcpyAsync(d_xyz, w->h_tri.data(), (size_t)nt * 9 * sizeof(float), cudaMemcpyHostToDevice, w->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_count, 0, (ncol + 1) * sizeof(uint32_t), w->stream);
    const uint32_t gt = (nt + RB_BLOCK - 1) / RB_BLOCK;
    if (e == cudaSuccess && nt) rb_tgrid_scatter_kernel<<<gt, RB_BLOCK, 0, w->stream>>>(G, d_xyz, nt, d_tri, d_count, nullptr, nullptr, nullptr);
    if (e == cudaSuccess) { rb_launch_scan(d_count, (uint32_t)ncol, d_start, d_tiles, w->stream); w->launches += 4; }
    uint32_t total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, d_start + ncol, 4, cudaMemcpyDeviceToHost, w->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
    if (e != cudaSuccess) { cleanup(); return fail(w, RB_ERR_CUDA, "static grid build: %s", cudaGetErrorString(e)); }
    uint32_t* d_list; float4* d_ent;
    { int r = dev_alloc(w, &d_list, std::max<size_t>(total, 1)); if (r) { cleanup(); return r; } }
    { int r = dev_alloc(w, &d_ent, std::max<size_t>(total, 1) * 2); if (r) { cleanup(); return r; } }
    e = cudaMemsetAsync(d_count, 0, (ncol + 1) * sizeof(uint32_t), w->stream);
    if (e == cudaSuccess && nt) rb_tgrid_scatter_kernel<<<gt, RB_BLOCK, 0, w->stream>>>(G, d_xyz, nt, nullptr, d_count, d_start, d_list, d_ent);
    if (e == cudaSuccess) rb_tgrid_finish_kernel<<<(uint32_t)((ncol + RB_BLOCK - 1) / RB_BLOCK), RB_BLOCK, 0, w->stream>>>((uint32_t)ncol, d_start, d_list, d_ent, d_yr);
    w->launches += 2;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
    cleanup();
    if (e != cudaSuccess) return fail(w, RB_ERR_CUDA, "static grid build: %s", cudaGetErrorString(e));
    P.tri = d_tri; P.tcol_start = d_start; P.tcol_tri = d_list; P.tcol_yr = d_yr; P.tcol_ent = d_ent;
    return RB_OK;
}

// GeometryMath::triangleCubeDetection for n (triangle, cube) pairs on the device: tri9 = n * (A, B, C), cube6 = n * (centre xyz,
// length, height, width); out[i] = 1 when they overlap
__global__ void __launch_bounds__(RB_BLOCK) rb_triangle_cube_kernel(uint32_t n, const float* __restrict__ tri9, const float* __restrict__ cube6, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* t = tri9 + 9ull * i; const float* q = cube6 + 6ull * i;
    out[i] = triangle_cube_detect(mk3(t[0], t[1], t[2]), mk3(t[3], t[4], t[5]), mk3(t[6], t[7], t[8]), mk3(q[0], q[1], q[2]), q[3], q[4], q[5]) ? 1u : 0u;
}
extern "C" int rb_triangle_cube_test(rb_world* w, uint32_t n, const float* tri9, const float* cube6, uint32_t* out) {
    if (!w || (n && (!tri9 || !cube6 || !out))) return w ? fail(w, RB_ERR_INVALID, "NULL argument") : RB_ERR_INVALID;
    if (!n) return RB_OK;
    CU(cudaSetDevice(w->cfg.device));
    float *d_t = nullptr, *d_c = nullptr; uint32_t* d_o = nullptr;
    CU(cudaMalloc(&d_t, (size_t)n * 36)); CU(cudaMalloc(&d_c, (size_t)n * 24)); CU(cudaMalloc(&d_o, (size_t)n * 4));
    cudaError_t e = cudaMemcpyAsync(d_t, tri9, (size_t)n * 36, cudaMemcpyHostToDevice, w->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_c, cube6, (size_t)n * 24, cudaMemcpyHostToDevice, w->stream);
    if (e == cudaSuccess) { rb_triangle_cube_kernel<<<(n + RB_BLOCK - 1) / RB_BLOCK, RB_BLOCK, 0, w->stream>>>(n, d_t, d_c, d_o); w->launches++; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, w->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
    cudaFree(d_t); cudaFree(d_c); cudaFree(d_o);
    if (e != cudaSuccess) return fail(w, RB_ERR_CUDA, "rb_triangle_cube_test: %s", cudaGetErrorString(e));
    return RB_OK;
}

// ---- collider ingest ------------------------------------------------------------------------------
// node transform = translation * rotX * rotY * rotZ * scale, trs9 = (tx,ty,tz, rx,ry,rz in degrees, sx,sy,sz)
int rb_ingest_trs_matrix(const float trs9[9], float out16[16]) {
    if (!trs9 || !out16) return RB_ERR_INVALID;
    M4 m = trs_matrix(trs9); memcpy(out16, m.m, sizeof m.m);
    return RB_OK;
}
// GeometryType::Triangle (FbxLoader.cpp:1133-1160): every index triple becomes a world-space triangle with the node TRS baked in
int rb_ingest_mesh_triangles(uint32_t nverts, const float* xyz3, uint32_t nidx, const uint32_t* idx, const float trs9[9], float* out_xyz9) {
    if ((nverts && !xyz3) || (nidx && (!idx || !out_xyz9)) || !trs9 || !idx_ok(nverts, nidx, idx)) return RB_ERR_INVALID;
    const M4 m = trs_matrix(trs9);
    for (uint32_t t = 0; t < nidx / 3; t++)
        for (int k = 0; k < 3; k++) {
            float q[4]; m_mulv(m, xyz3 + 3ull * idx[3 * t + k], 1.0f, q);
            out_xyz9[9ull * t + 3 * k] = q[0]; out_xyz9[9ull * t + 3 * k + 1] = q[1]; out_xyz9[9ull * t + 3 * k + 2] = q[2];
        }
    return RB_OK;
}
// GeometryType::Sphere (FbxLoader.cpp:1161-1215): ONE sphere per collider mesh, r = (max.x - min.x) / 2, centre = min + r on all
// three axes; max starts at the smallest positive float as in the reference. sphere4 = (cx, cy, cz, r): an object-space sphere
// for rb_add_model.
int rb_ingest_mesh_sphere(uint32_t nverts, const float* xyz3, uint32_t nidx, const uint32_t* idx, const float trs9[9], float sphere4[4]) {
    if ((nverts && !xyz3) || (nidx && !idx) || !trs9 || !sphere4 || !idx_ok(nverts, nidx, idx)) return RB_ERR_INVALID;
    const M4 m = trs_matrix(trs9);
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { FLT_MIN, FLT_MIN, FLT_MIN };
    for (uint32_t i = 0; i < nidx; i++) {
        float q[4]; m_mulv(m, xyz3 + 3ull * idx[i], 1.0f, q);
        for (int a = 0; a < 3; a++) { if (q[a] < mn[a]) mn[a] = q[a]; if (q[a] > mx[a]) mx[a] = q[a]; }
    }
    const float radius = (mx[0] - mn[0]) / 2.0f;
    sphere4[0] = mn[0] + radius; sphere4[1] = mn[1] + radius; sphere4[2] = mn[2] + radius; sphere4[3] = radius;
    return RB_OK;
}
// convenience: a static (GeometryType::Triangle) model straight into the world
int rb_ingest_static_mesh(rb_world* w, uint32_t nverts, const float* xyz3, uint32_t nidx, const uint32_t* idx, const float trs9[9], uint32_t* geom_id) {
    if (!w) return RB_ERR_INVALID;
    std::vector<float> tri((size_t)(nidx / 3) * 9);
    int r = rb_ingest_mesh_triangles(nverts, xyz3, nidx, idx, trs9, tri.data());
    if (r) return fail(w, r, "rb_ingest_static_mesh: bad mesh (NULL array or index out of range)");
    return rb_add_static_geometry(w, nidx / 3, tri.data(), geom_id);
}

} // extern "C"
