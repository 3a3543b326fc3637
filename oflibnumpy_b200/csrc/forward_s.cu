// Source-referenced (forward) resampling for sm_100a: replaces scipy.interpolate.griddata(points = grid + flow,
// values, grid, 'linear') + nan_to_num at utils.py:237-258 of the reference.
//
// griddata triangulates the displaced pixel positions (Qhull Delaunay) and interpolates barycentrically inside each
// triangle. For a non-folding field that triangulation is the displaced pixel grid itself with every cell split along
// its Delaunay diagonal, so the kernel pair below rasterises exactly that mesh:
//
//   fwd_scatter  one thread per source cell: positions in float64 (int + float32 flow, as numpy promotes), Delaunay
//                diagonal by the in-circle determinant, both triangles rasterised over their pixel bounding box with
//                canonically ordered edge functions (bit-identical on shared edges -> watertight, no cracks), each
//                covered pixel receives atomicMax(triangle id). The maximum is order independent, so the result is
//                deterministic; where a folding field covers a pixel more than once the largest source index wins.
//   fwd_gather   one thread per output pixel: decodes the winning triangle, recomputes its barycentric weights in
//                float64 and interpolates payload and mask; pixels no triangle covers are 0 / invalid (outside the
//                hull the reference returns NaN -> 0).
//
// Documented deviations: cells whose corners are exactly co-circular have no unique Delaunay diagonal
// (Qhull's choice there is an artefact of its merge order); with point_mask, cells touching a removed point are left
// empty instead of being bridged by long triangles; concave pockets between the displaced image border and its convex
// hull are not filled.
#include "ofk_common.cuh"

namespace ofk {

struct P2 {
    double x, y;
};

__device__ __forceinline__ P2 displaced(const float2* __restrict__ fl, int W, int row, int col, float sign) {
    const float2 f = __ldg(fl + row * W + col);
    P2 p;
    p.x = static_cast<double>(col) + static_cast<double>(sign * f.x);
    p.y = static_cast<double>(row) + static_cast<double>(sign * f.y);
    return p;
}

__device__ __forceinline__ double orient(const P2& a, const P2& b, const P2& c) {
    return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
}

// > 0 iff d lies inside the circumcircle of the positively oriented triangle (a, b, c)
__device__ __forceinline__ double incircle(const P2& a, const P2& b, const P2& c, const P2& d) {
    const double ax = a.x - d.x, ay = a.y - d.y, bx = b.x - d.x, by = b.y - d.y, cx = c.x - d.x, cy = c.y - d.y;
    const double a2 = ax * ax + ay * ay, b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
    return ax * (by * c2 - b2 * cy) - ay * (bx * c2 - b2 * cx) + a2 * (bx * cy - by * cx);
}

// Corner naming of cell (i, j): a = (i, j), b = (i, j+1), c = (i+1, j), d = (i+1, j+1).
// diag 0 splits along a-d: triangles (a, b, d) and (a, d, c); diag 1 along b-c: (a, b, c) and (b, d, c).
__device__ __forceinline__ int choose_diagonal(const P2& a, const P2& b, const P2& c, const P2& d) {
    const double o0a = orient(a, b, d), o0b = orient(a, d, c);   // diag 0 halves
    const double o1a = orient(a, b, c), o1b = orient(b, d, c);   // diag 1 halves
    const bool ok0 = (o0a > 0 && o0b > 0) || (o0a < 0 && o0b < 0);
    const bool ok1 = (o1a > 0 && o1b > 0) || (o1a < 0 && o1b < 0);
    if (ok0 && ok1) {  // convex quad: Delaunay criterion
        const double det = incircle(a, b, d, c);
        const bool c_inside = (o0a > 0) ? (det > 0) : (det < 0);
        return c_inside ? 1 : 0;
    }
    return ok1 && !ok0 ? 1 : 0;
}

// vertex k (0..2) of triangle `tri` (0/1) of a cell split by `diag`, as corner code 0=a 1=b 2=c 3=d
__device__ __forceinline__ int corner_of(int diag, int tri, int k) {
    // diag0: t0 = a b d, t1 = a d c ; diag1: t0 = a b c, t1 = b d c
    const int table = diag == 0 ? (tri == 0 ? 0x310 : 0x230) : (tri == 0 ? 0x210 : 0x231);
    return (table >> (4 * k)) & 0xf;
}

// Edge function of pixel q against edge (u, v), evaluated with the endpoints in canonical (index) order so the two
// triangles sharing an edge see bit-identical magnitudes with opposite signs.
__device__ __forceinline__ double edge_fn(const P2& u, int iu, const P2& v, int iv, double qx, double qy) {
    if (iu < iv) return (v.x - u.x) * (qy - u.y) - (v.y - u.y) * (qx - u.x);
    return -((u.x - v.x) * (qy - v.y) - (u.y - v.y) * (qx - v.x));
}

__global__ void __launch_bounds__(256, 4) fwd_scatter(const float* __restrict__ flow, float sign,
                                                   const uint8_t* __restrict__ point_mask,
                                                   unsigned int* __restrict__ winner, int H, int W) {
    const int j = blockIdx.x * 32 + (threadIdx.x & 31);
    const int i = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int n = blockIdx.z;
    if (i >= H - 1 || j >= W - 1) return;
    const size_t fbase = (size_t)n * H * W;
    const float2* fl = reinterpret_cast<const float2*>(flow) + fbase;
    if (point_mask != nullptr) {
        const uint8_t* pm = point_mask + fbase;
        const int o = i * W + j;
        if (!(pm[o] && pm[o + 1] && pm[o + W] && pm[o + W + 1])) return;
    }
    P2 v[4];
    v[0] = displaced(fl, W, i, j, sign);
    v[1] = displaced(fl, W, i, j + 1, sign);
    v[2] = displaced(fl, W, i + 1, j, sign);
    v[3] = displaced(fl, W, i + 1, j + 1, sign);
    const int vid[4] = {i * W + j, i * W + j + 1, (i + 1) * W + j, (i + 1) * W + j + 1};
    const int diag = choose_diagonal(v[0], v[1], v[2], v[3]);
    unsigned int* win = winner + fbase;
#pragma unroll
    for (int tri = 0; tri < 2; ++tri) {
        const int c0 = corner_of(diag, tri, 0), c1 = corner_of(diag, tri, 1), c2 = corner_of(diag, tri, 2);
        const P2 p0 = v[c0], p1 = v[c1], p2 = v[c2];
        const double area2 = orient(p0, p1, p2);
        if (area2 == 0.0) continue;
        const double sgn = area2 > 0 ? 1.0 : -1.0;
        // candidate pixels: bounding box in float32, rounded outwards (a superset is enough: coverage is decided by
        // the exact edge functions below)
        int xmin = (int)ceilf(fminf(__double2float_rd(p0.x), fminf(__double2float_rd(p1.x), __double2float_rd(p2.x))));
        int xmax = (int)floorf(fmaxf(__double2float_ru(p0.x), fmaxf(__double2float_ru(p1.x), __double2float_ru(p2.x))));
        int ymin = (int)ceilf(fminf(__double2float_rd(p0.y), fminf(__double2float_rd(p1.y), __double2float_rd(p2.y))));
        int ymax = (int)floorf(fmaxf(__double2float_ru(p0.y), fmaxf(__double2float_ru(p1.y), __double2float_ru(p2.y))));
        xmin = max(xmin, 0);
        ymin = max(ymin, 0);
        xmax = min(xmax, W - 1);
        ymax = min(ymax, H - 1);
        const unsigned int id = ((unsigned int)(i * W + j) << 2 | (unsigned int)(tri << 1) | (unsigned int)diag) + 1u;
        for (int y = ymin; y <= ymax; ++y) {
            for (int x = xmin; x <= xmax; ++x) {
                const double qx = x, qy = y;
                const double e0 = sgn * edge_fn(p1, vid[c1], p2, vid[c2], qx, qy);
                const double e1 = sgn * edge_fn(p2, vid[c2], p0, vid[c0], qx, qy);
                const double e2 = sgn * edge_fn(p0, vid[c0], p1, vid[c1], qx, qy);
                if (e0 >= 0.0 && e1 >= 0.0 && e2 >= 0.0) atomicMax(win + y * W + x, id);
            }
        }
    }
}

__global__ void __launch_bounds__(256) fwd_gather(const float* __restrict__ payload, int C,
                                                  const float* __restrict__ flow, float sign,
                                                  const uint8_t* __restrict__ payload_mask,
                                                  const unsigned int* __restrict__ winner, float* __restrict__ out,
                                                  uint8_t* __restrict__ out_mask, int rule, int H, int W,
                                                  unsigned long long inv_w /* ceil(2^64 / W) */) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int n = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t fbase = (size_t)n * H * W;
    const int pix = y * W + x;
    const unsigned int id = winner[fbase + pix];
    if (id == 0u) {
        for (int c = 0; c < C; ++c) out[(fbase + pix) * C + c] = 0.f;
        if (out_mask) out_mask[fbase + pix] = 0;
        return;
    }
    const unsigned int code = id - 1u;
    const int diag = code & 1, tri = (code >> 1) & 1, cell = (int)(code >> 2);
    const int i = (int)__umul64hi((unsigned long long)cell, inv_w), j = cell - i * W;   // exact: cell < 2^29
    const float2* fl = reinterpret_cast<const float2*>(flow) + fbase;
    int vidx[3];
    P2 p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int c = corner_of(diag, tri, k);
        const int r = i + (c >> 1), q = j + (c & 1);
        vidx[k] = r * W + q;
        p[k] = displaced(fl, W, r, q, sign);
    }
    const double area2 = orient(p[0], p[1], p[2]);
    const double qx = x, qy = y;
    // one reciprocal instead of two divisions: the weights move by an ulp at most (values are compared at 1e-3, the
    // mask tests below are decided by exact zeros and by margins far above an ulp)
    const double inv_area = __drcp_rn(area2);
    double w0 = edge_fn(p[1], vidx[1], p[2], vidx[2], qx, qy) * inv_area;
    double w1 = edge_fn(p[2], vidx[2], p[0], vidx[0], qx, qy) * inv_area;
    double w2 = 1.0 - w0 - w1;
    const float* pay = payload + fbase * C;
    for (int c = 0; c < C; ++c) {
        const double val = w0 * (double)__ldg(pay + (size_t)vidx[0] * C + c) +
                           w1 * (double)__ldg(pay + (size_t)vidx[1] * C + c) +
                           w2 * (double)__ldg(pay + (size_t)vidx[2] * C + c);
        out[(fbase + pix) * C + c] = (float)val;
    }
    if (out_mask) {
        double m = w0 + w1 + w2;
        if (payload_mask) {
            const uint8_t* pm = payload_mask + fbase;
            m = (pm[vidx[0]] ? w0 : 0.0) + (pm[vidx[1]] ? w1 : 0.0) + (pm[vidx[2]] ? w2 : 0.0);
        }
        bool ok;
        if (rule == OFK_RULE_STRICT) ok = ((float)m == 1.0f);   // float payloads: `== 1` after the float32 cast
        else ok = (m > 0.5);                                    // integer payloads: np.round(m) == 1
        out_mask[fbase + pix] = ok ? 1 : 0;
    }
}

}  // namespace ofk

using namespace ofk;

extern "C" size_t ofk_forward_s_workspace(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)N * H * W * sizeof(unsigned int);
}

extern "C" int ofk_forward_s(const float* payload, int C, const float* flow, float flow_sign,
                             const uint8_t* payload_mask, const uint8_t* point_mask, float* out, uint8_t* out_mask,
                             int mask_rule, int N, int H, int W, void* ws, size_t ws_bytes, ofk_stream_t stream) {
    OFK_CHECK_ARG(flow != nullptr, "ofk_forward_s: flow is NULL");
    OFK_CHECK_ARG(N >= 0 && H > 0 && W > 0, "ofk_forward_s: bad shape N=%d H=%d W=%d", N, H, W);
    OFK_CHECK_ARG(C >= 0 && (C == 0 || (payload != nullptr && out != nullptr)), "ofk_forward_s: payload/out NULL");
    OFK_CHECK_ARG(C > 0 || out_mask != nullptr, "ofk_forward_s: nothing to compute");
    OFK_CHECK_ARG(flow_sign == 1.0f || flow_sign == -1.0f, "ofk_forward_s: flow_sign must be +1 or -1");
    OFK_CHECK_ARG(mask_rule == OFK_RULE_STRICT || mask_rule == OFK_RULE_GT_HALF,
                  "ofk_forward_s: mask rule must be STRICT or GT_HALF");
    OFK_CHECK_ARG((size_t)H * W < ((size_t)1 << 29), "ofk_forward_s: frame too large for 32-bit triangle ids");
    OFK_CHECK_ARG((reinterpret_cast<uintptr_t>(flow) & 7) == 0, "ofk_forward_s: flow must be 8-byte aligned");
    if (N == 0) return OFK_OK;
    OFK_CHECK_ARG(N <= 65535, "ofk_forward_s: N=%d exceeds 65535", N);
    const size_t need = ofk_forward_s_workspace(N, H, W);
    OFK_CHECK_ARG(ws != nullptr && ws_bytes >= need, "ofk_forward_s: workspace of %zu bytes needed, got %zu", need,
                  ws_bytes);
    cudaStream_t st = as_stream(stream);
    unsigned int* winner = static_cast<unsigned int*>(ws);
    OFK_CUDA(cudaMemsetAsync(winner, 0, need, st));
    if (H > 1 && W > 1) {
        dim3 grid((W - 1 + 31) / 32, (H - 1 + 7) / 8, N);
        fwd_scatter<<<grid, 256, 0, st>>>(flow, flow_sign, point_mask, winner, H, W);
        OFK_LAUNCHED();
    }
    dim3 grid((W + 31) / 32, (H + 7) / 8, N);
    const unsigned long long inv_w = W > 1 ? ~0ull / (unsigned long long)W + 1ull : 0ull;   // ceil(2^64 / W), W >= 2
    fwd_gather<<<grid, 256, 0, st>>>(payload, C, flow, flow_sign, payload_mask, winner, out, out_mask, mask_rule, H, W,
                                     inv_w);
    OFK_LAUNCHED();
    return OFK_OK;
}

// ------------------------------------------------------------------------------------------- scattered -> scattered
// Barycentric interpolation of values attached to the displaced grid at arbitrary query points: replaces the direct
// `griddata(grid - A, A||mask, grid - B, 'linear', fill_value=0)` of combine_with mode 2 / ref 't'
// (flow_class.py:1398-1410) and the griddata calls of track_pts (utils.py:603,614). Same mesh as above (cells split
// along their Delaunay diagonal); the containing triangle of a query is found by a fixed-point walk
// p <- q - sign*flow(p) towards the source cell followed by an exact containment search around it.
namespace ofk {

__device__ __forceinline__ P2 mesh_vertex(const float2* __restrict__ fl, int W, int row, int col, float sign,
                                          int pos_f32) {
    const float2 f = __ldg(fl + row * W + col);
    P2 p;
    if (pos_f32) {  // the reference builds these coordinates in float32 (in-place adds on a float32 array)
        p.x = static_cast<double>(__fadd_rn(sign * f.x, static_cast<float>(col)));
        p.y = static_cast<double>(__fadd_rn(sign * f.y, static_cast<float>(row)));
    } else {
        p.x = static_cast<double>(col) + static_cast<double>(sign * f.x);
        p.y = static_cast<double>(row) + static_cast<double>(sign * f.y);
    }
    return p;
}

// tests cell (i, j); on success fills the three vertex indices and barycentric weights
__device__ bool locate_in_cell(const float2* __restrict__ fl, int H, int W, float sign, int pos_f32, int i, int j,
                               double qx, double qy, int (&vidx)[3], double (&w)[3]) {
    if (i < 0 || j < 0 || i >= H - 1 || j >= W - 1) return false;
    P2 v[4];
    v[0] = mesh_vertex(fl, W, i, j, sign, pos_f32);
    v[1] = mesh_vertex(fl, W, i, j + 1, sign, pos_f32);
    v[2] = mesh_vertex(fl, W, i + 1, j, sign, pos_f32);
    v[3] = mesh_vertex(fl, W, i + 1, j + 1, sign, pos_f32);
    const int vid[4] = {i * W + j, i * W + j + 1, (i + 1) * W + j, (i + 1) * W + j + 1};
    const int diag = choose_diagonal(v[0], v[1], v[2], v[3]);
    for (int tri = 0; tri < 2; ++tri) {
        const int c0 = corner_of(diag, tri, 0), c1 = corner_of(diag, tri, 1), c2 = corner_of(diag, tri, 2);
        const double area2 = orient(v[c0], v[c1], v[c2]);
        if (area2 == 0.0) continue;
        const double sgn = area2 > 0 ? 1.0 : -1.0;
        const double e0 = edge_fn(v[c1], vid[c1], v[c2], vid[c2], qx, qy);
        const double e1 = edge_fn(v[c2], vid[c2], v[c0], vid[c0], qx, qy);
        const double e2 = edge_fn(v[c0], vid[c0], v[c1], vid[c1], qx, qy);
        if (sgn * e0 >= 0.0 && sgn * e1 >= 0.0 && sgn * e2 >= 0.0) {
            vidx[0] = vid[c0];
            vidx[1] = vid[c1];
            vidx[2] = vid[c2];
            w[0] = e0 / area2;
            w[1] = e1 / area2;
            w[2] = 1.0 - w[0] - w[1];
            return true;
        }
    }
    return false;
}

__global__ void __launch_bounds__(128) mesh_sample_kernel(const float* __restrict__ mesh_flow, float mesh_sign,
                                                          int pos_f32, const float* __restrict__ payload, int C,
                                                          const uint8_t* __restrict__ payload_mask,
                                                          const float* __restrict__ query_flow, float query_sign,
                                                          const double* __restrict__ query_pts, int Q,
                                                          float* __restrict__ out, float* __restrict__ out_maskval,
                                                          uint8_t* __restrict__ found, int H, int W) {
    const int n = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Q) return;
    const size_t fbase = (size_t)n * H * W;
    const float2* fl = reinterpret_cast<const float2*>(mesh_flow) + fbase;
    double qx, qy;
    if (query_pts != nullptr) {  // (row, col) pairs
        qy = query_pts[((size_t)n * Q + k) * 2];
        qx = query_pts[((size_t)n * Q + k) * 2 + 1];
    } else {                     // one query per pixel: p + query_sign * query_flow[p]
        const int row = k / W, col = k - row * W;
        const float2 f = __ldg(reinterpret_cast<const float2*>(query_flow) + fbase + k);
        if (pos_f32) {
            qx = static_cast<double>(__fadd_rn(query_sign * f.x, static_cast<float>(col)));
            qy = static_cast<double>(__fadd_rn(query_sign * f.y, static_cast<float>(row)));
        } else {
            qx = static_cast<double>(col) + static_cast<double>(query_sign * f.x);
            qy = static_cast<double>(row) + static_cast<double>(query_sign * f.y);
        }
    }
    // fixed-point walk towards the source position p with p + sign*flow(p) = q (bilinear flow lookup, clamped)
    double px = qx, py = qy;
    for (int it = 0; it < 12; ++it) {
        const double cx = fmin(fmax(px, 0.0), (double)(W - 1)), cy = fmin(fmax(py, 0.0), (double)(H - 1));
        const int j0 = min((int)cx, W - 2 < 0 ? 0 : W - 2), i0 = min((int)cy, H - 2 < 0 ? 0 : H - 2);
        const double a = cx - j0, b = cy - i0;
        const int j1 = min(j0 + 1, W - 1), i1 = min(i0 + 1, H - 1);
        const float2 f00 = __ldg(fl + i0 * W + j0), f01 = __ldg(fl + i0 * W + j1), f10 = __ldg(fl + i1 * W + j0),
                     f11 = __ldg(fl + i1 * W + j1);
        const double u = (1 - b) * ((1 - a) * f00.x + a * f01.x) + b * ((1 - a) * f10.x + a * f11.x);
        const double v = (1 - b) * ((1 - a) * f00.y + a * f01.y) + b * ((1 - a) * f10.y + a * f11.y);
        const double nx = qx - mesh_sign * u, ny = qy - mesh_sign * v;
        const bool done = fabs(nx - px) < 1e-3 && fabs(ny - py) < 1e-3;
        px = nx;
        py = ny;
        if (done) break;
    }
    int vidx[3];
    double w[3];
    bool hit = false;
    const int ci = (int)floor(py), cj = (int)floor(px);
    for (int ring = 0; ring <= 2 && !hit; ++ring) {
        for (int di = -ring; di <= ring && !hit; ++di)
            for (int dj = -ring; dj <= ring && !hit; ++dj) {
                if (max(abs(di), abs(dj)) != ring) continue;
                hit = locate_in_cell(fl, H, W, mesh_sign, pos_f32, ci + di, cj + dj, qx, qy, vidx, w);
            }
    }
    const size_t o = (size_t)n * Q + k;
    if (found) found[o] = hit ? 1 : 0;
    if (!hit) {
        for (int c = 0; c < C; ++c) out[o * C + c] = 0.f;
        if (out_maskval) out_maskval[o] = 0.f;
        return;
    }
    const float* pay = payload + fbase * C;
    for (int c = 0; c < C; ++c) {
        const double val = w[0] * (double)__ldg(pay + (size_t)vidx[0] * C + c) +
                           w[1] * (double)__ldg(pay + (size_t)vidx[1] * C + c) +
                           w[2] * (double)__ldg(pay + (size_t)vidx[2] * C + c);
        out[o * C + c] = (float)val;
    }
    if (out_maskval) {
        double m = w[0] + w[1] + w[2];
        if (payload_mask) {
            const uint8_t* pm = payload_mask + fbase;
            m = (pm[vidx[0]] ? w[0] : 0.0) + (pm[vidx[1]] ? w[1] : 0.0) + (pm[vidx[2]] ? w[2] : 0.0);
        }
        out_maskval[o] = (float)m;
    }
}

}  // namespace ofk

extern "C" int ofk_mesh_sample(const float* mesh_flow, float mesh_sign, int pos_f32, const float* payload, int C,
                               const uint8_t* payload_mask, const float* query_flow, float query_sign,
                               const double* query_pts, int Q, float* out, float* out_maskval, uint8_t* found, int N,
                               int H, int W, ofk_stream_t stream) {
    OFK_CHECK_ARG(mesh_flow != nullptr, "ofk_mesh_sample: mesh_flow is NULL");
    OFK_CHECK_ARG((query_flow != nullptr) != (query_pts != nullptr),
                  "ofk_mesh_sample: exactly one of query_flow / query_pts must be given");
    OFK_CHECK_ARG(N >= 0 && H > 1 && W > 1 && C >= 0, "ofk_mesh_sample: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
    OFK_CHECK_ARG(C == 0 || (payload != nullptr && out != nullptr), "ofk_mesh_sample: payload/out NULL");
    OFK_CHECK_ARG((size_t)H * W < ((size_t)1 << 30), "ofk_mesh_sample: frame too large");
    if (query_flow != nullptr) Q = H * W;
    OFK_CHECK_ARG(Q >= 0, "ofk_mesh_sample: negative query count");
    if (N == 0 || Q == 0) return OFK_OK;
    OFK_CHECK_ARG(N <= 65535, "ofk_mesh_sample: N too large");
    dim3 grid((Q + 127) / 128, N);
    ofk::mesh_sample_kernel<<<grid, 128, 0, ofk::as_stream(stream)>>>(mesh_flow, mesh_sign, pos_f32, payload, C,
                                                                     payload_mask, query_flow, query_sign, query_pts,
                                                                     Q, out, out_maskval, found, H, W);
    OFK_LAUNCHED();
    return OFK_OK;
}
