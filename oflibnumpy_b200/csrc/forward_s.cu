// Source-referenced (forward) resampling for sm_100a: replaces scipy.interpolate.griddata(points = grid + flow,
// values, grid, 'linear') + nan_to_num at utils.py:237-258 of the reference (Flow.apply ref 's', same-reference
// invert, switch_ref, combine modes 1 / 2, the 's'-side valid areas).
//
// griddata triangulates the displaced pixel positions (Qhull Delaunay, AFTER dropping the points a `consider_mask`
// removes) and interpolates barycentrically inside each triangle, 0 outside the convex hull. The algorithm and its
// predicates live in forward_geom.cuh / forward_irregular.cuh (shared with the host build the CPU tests run against
// the reference's goldens); this file maps it onto the GPU:
//
//   fwd_raster_kernel   regular part. A CTA owns a tile of 64 x 16 source cells: the 65 x 17 displaced vertices
//                       (float64 positions, computed once per vertex), their payload and mask bytes are staged in
//                       shared memory; one thread per cell column walks 4 rows keeping the shared row of vertices in
//                       registers, picks the Delaunay diagonal and rasterises both triangles. The fill rule makes every
//                       pixel inside the mesh the product of exactly one triangle, so values and validity go straight
//                       to the output: no triangle-id plane in HBM, no atomics, no second pass over the frame.
//   irr_*               irregular part: pixels no intact cell produced (marker byte left in the validity plane) are
//                       located in the Delaunay triangulation of the boundary sites (frame border, rims of removed
//                       points), binned by position with two counting passes and a scan; a per-frame hull polygon
//                       rejects the pixels outside the hull without a search.
//   legacy::fwd_scatter / fwd_gather   folding fields (a displaced cell with a non-positive triangle): several
//                       triangles cover a pixel, "exactly one" no longer holds; such frames are redone completely by
//                       the order-independent atomicMax(triangle id) resolve of round 1 (largest source index wins).
//
// Deviation left: cells whose four corners are co-circular to within rounding have no unique Delaunay
// diagonal (similarity transforms of the pixel grid); Qhull's pick there is an artefact of its facet merging.
#include "ofk_common.cuh"
#include "forward_irregular.cuh"

// ============================================================ order-independent resolve for folding fields (round 1)
namespace ofk {
namespace legacy {

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
                                                   unsigned int* __restrict__ winner, int H, int W,
                                                   const int* __restrict__ folded) {
    const int j = blockIdx.x * 32 + (threadIdx.x & 31);
    const int i = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int n = blockIdx.z;
    if (!folded[n]) return;
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
                                                  unsigned long long inv_w /* ceil(2^64 / W) */,
                                                  const int* __restrict__ folded) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int n = blockIdx.z;
    if (!folded[n]) return;
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


__global__ void __launch_bounds__(256) fwd_clear(unsigned int* __restrict__ winner, size_t frame_px,
                                                 const int* __restrict__ folded) {
    const int n = blockIdx.y;
    if (!folded[n]) return;
    const size_t t = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
    for (int k = 0; k < 4; ++k)
        if (t + k < frame_px) winner[(size_t)n * frame_px + t + k] = 0u;
}

}  // namespace legacy
}  // namespace ofk

// =========================================================================================== the B200 pipeline
namespace ofk {
namespace fwdk {
using namespace fwd;

__device__ unsigned long long g_stats[4];   // located, outside (by search), failed walks, rejected by the hull filter

constexpr int TW = 64, TH = 16, SW = TW + 1, SH = TH + 1, NV = SW * SH;
constexpr uint8_t UNCOVERED = 0xFF;

struct RasterArgs {
    const float* payload;
    const float* flow;
    const uint8_t* payload_mask;
    const uint8_t* point_mask;
    float* out;
    uint8_t* out_mask;
    uint8_t* cover;      // == out_mask when the caller wants the validity plane, else a workspace plane
    int* folded;
    float sign;
    int C, rule_strict, H, W;
    double flip_tol;     // test hook, 0 in production (see cell_diagonal)
};

template <int CT>
struct EmitDev {
    const RasterArgs& A;
    const float* s_pay;
    const uint8_t* s_pm;
    int sbase;           // shared-memory index of corner a
    size_t frame;        // n * H * W
    int gi, gj;          // cell
    __device__ __forceinline__ void operator()(int x, int y, int k0, int k1, int k2, double w0, double w1,
                                               double w2) const {
        const int v0 = sbase + (k0 >> 1) * SW + (k0 & 1), v1 = sbase + (k1 >> 1) * SW + (k1 & 1),
                  v2 = sbase + (k2 >> 1) * SW + (k2 & 1);
        const size_t px = frame + (size_t)y * A.W + x;
        const bool m0 = s_pm[v0] != 0, m1 = s_pm[v1] != 0, m2 = s_pm[v2] != 0;
        if (CT >= 0) {
            constexpr int C = CT > 0 ? CT : 0;
            interp_store(s_pay + v0 * C, s_pay + v1 * C, s_pay + v2 * C, m0, m1, m2, w0, w1, w2, C, A.out + px * C,
                         A.out_mask ? A.out_mask + px : nullptr, A.rule_strict);
        } else {
            const size_t g0 = frame + (size_t)(gi + (k0 >> 1)) * A.W + gj + (k0 & 1),
                         g1 = frame + (size_t)(gi + (k1 >> 1)) * A.W + gj + (k1 & 1),
                         g2 = frame + (size_t)(gi + (k2 >> 1)) * A.W + gj + (k2 & 1);
            interp_store(A.payload + g0 * A.C, A.payload + g1 * A.C, A.payload + g2 * A.C, m0, m1, m2, w0, w1, w2, A.C,
                         A.out + px * A.C, A.out_mask ? A.out_mask + px : nullptr, A.rule_strict);
        }
        if (A.out_mask == nullptr) A.cover[px] = 1;
    }
};

template <int CT>
__global__ void __launch_bounds__(256) fwd_raster_kernel(const RasterArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int PC = CT > 0 ? CT : 0;
    P2* s_pos = reinterpret_cast<P2*>(smem);
    float* s_pay = reinterpret_cast<float*>(s_pos + NV);
    uint8_t* s_pm = reinterpret_cast<uint8_t*>(s_pay + NV * PC);
    uint8_t* s_pt = s_pm + NV;
    const int n = blockIdx.z, i0 = blockIdx.y * TH, j0 = blockIdx.x * TW;
    const size_t frame = (size_t)n * A.H * A.W;
    const float2* fl = reinterpret_cast<const float2*>(A.flow) + frame;
    for (int v = threadIdx.x; v < NV; v += 256) {
        const int r = v / SW, c = v - r * SW;
        const int gi = min(i0 + r, A.H - 1), gj = min(j0 + c, A.W - 1);
        const size_t g = (size_t)gi * A.W + gj;
        const float2 f = __ldg(fl + g);
        s_pos[v] = displaced(f.x, f.y, gi, gj, A.sign);
#pragma unroll
        for (int k = 0; k < PC; ++k) s_pay[v * PC + k] = __ldg(A.payload + (frame + g) * PC + k);
        s_pm[v] = A.payload_mask ? A.payload_mask[frame + g] : (uint8_t)1;
        s_pt[v] = A.point_mask ? A.point_mask[frame + g] : (uint8_t)1;
    }
    __syncthreads();
    const int jl = threadIdx.x & (TW - 1), strip = threadIdx.x / TW;
    const int gj = j0 + jl;
    if (gj >= A.W - 1) return;
    constexpr int ROWS = TH / (256 / TW);
    int il = strip * ROWS;
    P2 a = s_pos[il * SW + jl], b = s_pos[il * SW + jl + 1];
    bool vab = s_pt[il * SW + jl] && s_pt[il * SW + jl + 1];
    for (int k = 0; k < ROWS; ++k, ++il) {
        const int gi = i0 + il;
        if (gi >= A.H - 1) break;
        const int sb = il * SW + jl;
        const P2 c = s_pos[sb + SW], d = s_pos[sb + SW + 1];
        const bool vcd = s_pt[sb + SW] && s_pt[sb + SW + 1];
        if (vab && vcd) {
            double area2[2];
            const int diag = cell_diagonal(a, b, c, d, area2, A.flip_tol);
            if (diag < 0) {
                A.folded[n] = 1;
            } else {
                EmitDev<CT> e{A, s_pay, s_pm, sb, frame, gi, gj};
                raster_cell(a, b, c, d, diag, area2, A.W, A.H, e);
            }
        }
        a = c;
        b = d;
        vab = vcd;
    }
}

// ------------------------------------------------------------------------------------------- boundary sites -> bins
struct IrrArgs {
    const float* flow;
    const uint8_t* point_mask;
    const int* folded;
    uint32_t* bins;      // [N][nb + 1]
    uint32_t* coarse;    // [N][nc]
    uint32_t* sites;     // [N][H * W]
    float sign;
    int H, W, nbx, nby, ncx, ncy;
    int perimeter_only;  // no point mask: the boundary sites are the frame border
};

__device__ __forceinline__ bool irr_site_of_thread(const IrrArgs& A, int n, long long t, int& row, int& col) {
    if (A.perimeter_only) {
        const long long P = 2ll * A.W + 2ll * (A.H - 2);
        if (t >= P) return false;
        if (t < A.W) {
            row = 0;
            col = (int)t;
        } else if (t < 2ll * A.W) {
            row = A.H - 1;
            col = (int)(t - A.W);
        } else {
            const long long u = t - 2ll * A.W;
            row = 1 + (int)(u >> 1);
            col = (u & 1) ? A.W - 1 : 0;
        }
        return true;
    }
    if (t >= (long long)A.H * A.W) return false;
    row = (int)(t / A.W);
    col = (int)(t - (long long)row * A.W);
    return is_boundary_site(A.point_mask ? A.point_mask + (size_t)n * A.H * A.W : nullptr, A.H, A.W, row, col);
}

template <int PASS>   // 0: count sites per bin, 1: fill the site lists (bins hold the end offsets, counted down)
__global__ void __launch_bounds__(256) irr_sites_kernel(const IrrArgs A) {
    const int n = blockIdx.y;
    if (A.folded[n]) return;
    int row, col;
    if (!irr_site_of_thread(A, n, (long long)blockIdx.x * 256 + threadIdx.x, row, col)) return;
    const size_t frame = (size_t)n * A.H * A.W;
    const uint32_t id = (uint32_t)(row * A.W + col);
    const float2 f = __ldg(reinterpret_cast<const float2*>(A.flow) + frame + id);
    const P2 p = displaced(f.x, f.y, row, col, A.sign);
    const int bx = bin_coord(p.x, A.nbx), by = bin_coord(p.y, A.nby);
    const int nb = A.nbx * A.nby;
    uint32_t* bins = A.bins + (size_t)n * (nb + 1);
    if (PASS == 0) {
        atomicAdd(bins + by * A.nbx + bx, 1u);
        atomicAdd(A.coarse + (size_t)n * A.ncx * A.ncy + (by >> COARSE_SHIFT) * A.ncx + (bx >> COARSE_SHIFT), 1u);
    } else {
        const uint32_t slot = atomicSub(bins + by * A.nbx + bx, 1u) - 1u;
        A.sites[frame + slot] = id;
    }
}

// counts -> inclusive end offsets, one CTA per frame; bins[nb] = number of sites
__global__ void __launch_bounds__(1024) irr_scan_kernel(uint32_t* bins_all, int nb) {
    uint32_t* a = bins_all + (size_t)blockIdx.x * (nb + 1);
    const int per = (nb + 1023) / 1024;
    const int b0 = min(threadIdx.x * per, nb), b1 = min(b0 + per, nb);
    uint32_t sum = 0;
    for (int b = b0; b < b1; ++b) sum += a[b];
    __shared__ uint32_t part[1024];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
    for (int b = b0; b < b1; ++b) {
        run += a[b];
        a[b] = run;
    }
    if (threadIdx.x == 1023) a[nb] = part[1023];
}

// ------------------------------------------------------------------------------------------------- hull pre-filter
struct HullWs {   // per frame
    unsigned long long dotkey[HULL_DIRS];
    unsigned long long slackkey[HULL_DIRS];
    uint32_t ext[HULL_DIRS];
};

// order-preserving map double -> uint64 (for atomicMax / atomicMin on floating-point values)
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double order_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

struct HullArgs {
    const float* flow;
    const int* folded;
    const uint32_t* bins;
    const uint32_t* sites;
    HullWs* ws;
    HullInfo* info;
    HullDirs dirs;
    float sign;
    int H, W, nb;
};

__device__ __forceinline__ SiteGrid hull_grid(const HullArgs& A, int n) {
    SiteGrid g;
    g.H = A.H;
    g.W = A.W;
    g.nbx = g.nby = g.ncx = g.ncy = 0;
    g.bin_start = nullptr;
    g.coarse = nullptr;
    g.sites = A.sites + (size_t)n * A.H * A.W;
    g.flow = A.flow + 2 * (size_t)n * A.H * A.W;
    g.sign = A.sign;
    return g;
}

constexpr int HULL_ITEMS = 4;   // sites per thread

template <int PASS>   // 0: largest dot product per direction, 1: smallest site id among the maximisers, 2: edge slack
__global__ void __launch_bounds__(256) hull_sites_kernel(const HullArgs A) {
    const int n = blockIdx.y;
    if (A.folded[n]) return;
    const uint32_t total = A.bins[(size_t)n * (A.nb + 1) + A.nb];
    const uint32_t s0 = blockIdx.x * (256 * HULL_ITEMS);
    if (s0 >= total) return;
    const SiteGrid g = hull_grid(A, n);
    HullWs& ws = A.ws[n];
    const HullInfo& info = A.info[n];
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < HULL_ITEMS; ++it) {
        const uint32_t s = s0 + it * 256 + threadIdx.x;
        const bool have = s < total;
        const uint32_t id = have ? g.sites[s] : 0u;
        P2 p;
        p.x = p.y = 0.0;
        if (have) p = site_pos(g, id);
        const int count = PASS == 2 ? info.m : HULL_DIRS;
        for (int k = 0; k < count; ++k) {
            if (PASS == 0) {
                unsigned long long key = have ? order_key(dfma(A.dirs.dx[k], p.x, dmul(A.dirs.dy[k], p.y))) : 0ull;
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other > key ? other : key;
                }
                if (lane == 0 && key != 0ull) atomicMax(&ws.dotkey[k], key);
            } else if (PASS == 1) {
                const bool hit = have && order_key(dfma(A.dirs.dx[k], p.x, dmul(A.dirs.dy[k], p.y))) == ws.dotkey[k];
                if (hit) atomicMin(&ws.ext[k], id);
            } else {
                unsigned long long key = have ? order_key(hull_edge_orient(info, k, p)) : ~0ull;
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other < key ? other : key;
                }
                if (lane == 0) atomicMin(&ws.slackkey[k], key);
            }
        }
    }
}

template <int PASS>   // 0: polygon of the extreme sites, slack keys reset; 1: slack keys -> HullInfo
__global__ void hull_frame_kernel(const HullArgs A) {
    const int n = blockIdx.x;
    if (A.folded[n]) return;
    HullWs& ws = A.ws[n];
    HullInfo& info = A.info[n];
    if (PASS == 0) {
        if (threadIdx.x == 0) {
            const SiteGrid g = hull_grid(A, n);
            uint32_t ext[HULL_DIRS];
            for (int k = 0; k < HULL_DIRS; ++k) ext[k] = ws.ext[k];
            hull_polygon(g, ext, info);
        }
        if (threadIdx.x < HULL_DIRS) ws.slackkey[threadIdx.x] = order_key(0.0);
    } else if (threadIdx.x < HULL_DIRS) {
        info.slack[threadIdx.x] = order_value(ws.slackkey[threadIdx.x]);
    }
}

// ------------------------------------------------------------------------------------------------- uncovered pixels
struct SolveArgs {
    const float* payload;
    const float* flow;
    const uint8_t* payload_mask;
    const int* folded;
    const uint32_t* bins;
    const uint32_t* coarse;
    const uint32_t* sites;
    const HullInfo* info;
    float* out;
    uint8_t* out_mask;
    const uint8_t* cover;
    float sign;
    int C, rule_strict, H, W, nbx, nby, ncx, ncy;
};

__global__ void __launch_bounds__(256) irr_solve_kernel(const SolveArgs A) {
    const int n = blockIdx.z;
    if (A.folded[n]) return;
    __shared__ HullInfo s_hull;
    {
        const int words = sizeof(HullInfo) / 4;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(A.info + n);
        for (int k = threadIdx.x; k < words; k += 256) reinterpret_cast<uint32_t*>(&s_hull)[k] = src[k];
    }
    __syncthreads();
    const int y = blockIdx.y;
    const size_t frame = (size_t)n * A.H * A.W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xbase = (blockIdx.x * 8 + warp) * 128;
    SiteGrid g;
    g.H = A.H;
    g.W = A.W;
    g.nbx = A.nbx;
    g.nby = A.nby;
    g.ncx = A.ncx;
    g.ncy = A.ncy;
    const int nb = A.nbx * A.nby;
    g.bin_start = A.bins + (size_t)n * (nb + 1);
    g.coarse = A.coarse + (size_t)n * A.ncx * A.ncy;
    g.sites = A.sites + frame;
    g.flow = A.flow + 2 * frame;
    g.sign = A.sign;
    for (int k = 0; k < 4; ++k) {
        const int x = xbase + k * 32 + lane;
        if (x >= A.W) break;
        const size_t px = frame + (size_t)y * A.W + x;
        if (A.cover[px] != UNCOVERED) continue;
        P2 q;
        q.x = x;
        q.y = y;
        uint32_t vid[3];
        double w[3];
        int st;
        if (hull_rejects(s_hull, q)) {
            st = LOC_OUTSIDE;
            atomicAdd(&g_stats[3], 1ull);
        } else {
            st = locate(g, q, vid, w);
            atomicAdd(&g_stats[st == LOC_FOUND ? 0 : (st == LOC_OUTSIDE ? 1 : 2)], 1ull);
        }
        if (st == LOC_FOUND) {
            const uint8_t* pm = A.payload_mask ? A.payload_mask + frame : nullptr;
            const float* pay = A.payload + frame * A.C;
            interp_store(pay + (size_t)vid[0] * A.C, pay + (size_t)vid[1] * A.C, pay + (size_t)vid[2] * A.C,
                         pm ? pm[vid[0]] != 0 : true, pm ? pm[vid[1]] != 0 : true, pm ? pm[vid[2]] != 0 : true, w[0],
                         w[1], w[2], A.C, A.out + px * A.C, A.out_mask ? A.out_mask + px : nullptr, A.rule_strict);
        } else {
            for (int c = 0; c < A.C; ++c) A.out[px * A.C + c] = 0.f;
            if (A.out_mask) A.out_mask[px] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------- workspace layout
struct WsLayout {
    size_t sites, cover, bins, coarse, hullws, hullinfo, folded, total;
    size_t zero_begin, zero_bytes;   // region cleared before every call (bins, coarse, hull keys, folded flags)
    int nbx, nby, ncx, ncy, nb, nc;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static WsLayout ws_layout(int N, int H, int W) {
    WsLayout L;
    L.nbx = grid_bins(W);
    L.nby = grid_bins(H);
    L.ncx = grid_coarse(L.nbx);
    L.ncy = grid_coarse(L.nby);
    L.nb = L.nbx * L.nby;
    L.nc = L.ncx * L.ncy;
    const size_t px = (size_t)N * H * W;
    size_t o = 0;
    L.sites = o;            // also the triangle-id plane of the legacy resolve (never used at the same time)
    o = align_up(o + px * 4, 256);
    L.cover = o;
    o = align_up(o + px, 256);
    L.hullinfo = o;
    o = align_up(o + (size_t)N * sizeof(HullInfo), 256);
    L.zero_begin = o;
    L.bins = o;
    o = align_up(o + (size_t)N * (L.nb + 1) * 4, 256);
    L.coarse = o;
    o = align_up(o + (size_t)N * L.nc * 4, 256);
    L.hullws = o;
    o = align_up(o + (size_t)N * sizeof(HullWs), 256);
    L.folded = o;
    o = align_up(o + (size_t)N * 4, 256);
    L.zero_bytes = o - L.zero_begin;
    L.total = o;
    return L;
}

__global__ void hull_ws_init_kernel(HullWs* ws, int N) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N * HULL_DIRS) ws[t / HULL_DIRS].ext[t % HULL_DIRS] = NO_SITE;
}

unsigned long long stat(int which) {
    unsigned long long v[4] = {0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(v, g_stats, sizeof(v)) != cudaSuccess) cudaGetLastError();
    return which >= 0 && which < 4 ? v[which] : 0ull;
}

}  // namespace fwdk

unsigned long long forward_s_stat(int which) { return fwdk::stat(which); }
static std::atomic<double> g_flip_tol{0.0};

}  // namespace ofk

using namespace ofk;

extern "C" int ofk_forward_s_set_flip_tol(double tol) {
    OFK_CHECK_ARG(tol >= 0.0, "ofk_forward_s_set_flip_tol: negative tolerance");
    g_flip_tol.store(tol);
    return OFK_OK;
}

extern "C" size_t ofk_forward_s_workspace(int N, int H, int W) {
    if (N <= 0 || H <= 0 || W <= 0) return 0;
    return fwdk::ws_layout(N, H, W).total;
}

template <int CT>
static void launch_raster(const fwdk::RasterArgs& A, int N, cudaStream_t st) {
    using namespace fwdk;
    constexpr int PC = CT > 0 ? CT : 0;
    const size_t smem = (size_t)NV * (sizeof(fwd::P2) + 4 * PC + 2);
    dim3 grid((A.W - 1 + TW - 1) / TW, (A.H - 1 + TH - 1) / TH, N);
    fwd_raster_kernel<CT><<<grid, 256, smem, st>>>(A);
}

extern "C" int ofk_forward_s(const float* payload, int C, const float* flow, float flow_sign,
                             const uint8_t* payload_mask, const uint8_t* point_mask, float* out, uint8_t* out_mask,
                             int mask_rule, int N, int H, int W, void* ws, size_t ws_bytes, ofk_stream_t stream) {
    using namespace fwdk;
    OFK_CHECK_ARG(flow != nullptr, "ofk_forward_s: flow is NULL");
    OFK_CHECK_ARG(N >= 0 && H > 0 && W > 0, "ofk_forward_s: bad shape N=%d H=%d W=%d", N, H, W);
    OFK_CHECK_ARG(C >= 0 && (C == 0 || (payload != nullptr && out != nullptr)), "ofk_forward_s: payload/out NULL");
    OFK_CHECK_ARG(C > 0 || out_mask != nullptr, "ofk_forward_s: nothing to compute");
    OFK_CHECK_ARG(flow_sign == 1.0f || flow_sign == -1.0f, "ofk_forward_s: flow_sign must be +1 or -1");
    OFK_CHECK_ARG(mask_rule == OFK_RULE_STRICT || mask_rule == OFK_RULE_GT_HALF,
                  "ofk_forward_s: mask rule must be STRICT or GT_HALF");
    OFK_CHECK_ARG((size_t)H * W < ((size_t)1 << 29), "ofk_forward_s: frame too large for 32-bit triangle ids");
    OFK_CHECK_ARG(H <= 65535, "ofk_forward_s: H=%d exceeds 65535", H);
    OFK_CHECK_ARG((reinterpret_cast<uintptr_t>(flow) & 7) == 0, "ofk_forward_s: flow must be 8-byte aligned");
    if (N == 0) return OFK_OK;
    OFK_CHECK_ARG(N <= 65535, "ofk_forward_s: N=%d exceeds 65535", N);
    const WsLayout L = ws_layout(N, H, W);
    OFK_CHECK_ARG(ws != nullptr && ws_bytes >= L.total, "ofk_forward_s: workspace of %zu bytes needed, got %zu", L.total,
                  ws_bytes);
    cudaStream_t st = as_stream(stream);
    unsigned char* base = static_cast<unsigned char*>(ws);
    uint32_t* d_sites = reinterpret_cast<uint32_t*>(base + L.sites);
    uint8_t* d_cover = out_mask != nullptr ? out_mask : reinterpret_cast<uint8_t*>(base + L.cover);
    uint32_t* d_bins = reinterpret_cast<uint32_t*>(base + L.bins);
    uint32_t* d_coarse = reinterpret_cast<uint32_t*>(base + L.coarse);
    HullWs* d_hullws = reinterpret_cast<HullWs*>(base + L.hullws);
    fwd::HullInfo* d_info = reinterpret_cast<fwd::HullInfo*>(base + L.hullinfo);
    int* d_folded = reinterpret_cast<int*>(base + L.folded);
    const size_t px = (size_t)N * H * W;
    const int strict = mask_rule == OFK_RULE_STRICT ? 1 : 0;

    OFK_CUDA(cudaMemsetAsync(base + L.zero_begin, 0, L.zero_bytes, st));
    OFK_CUDA(cudaMemsetAsync(d_cover, UNCOVERED, px, st));
    hull_ws_init_kernel<<<(N * fwd::HULL_DIRS + 255) / 256, 256, 0, st>>>(d_hullws, N);
    OFK_LAUNCHED();

    // ---- regular part
    if (H > 1 && W > 1) {
        RasterArgs A{payload, flow, payload_mask, point_mask, out, out_mask, d_cover, d_folded, flow_sign, C, strict,
                     H, W, g_flip_tol.load()};
        switch (C) {
            case 0: launch_raster<0>(A, N, st); break;
            case 1: launch_raster<1>(A, N, st); break;
            case 2: launch_raster<2>(A, N, st); break;
            case 3: launch_raster<3>(A, N, st); break;
            case 4: launch_raster<4>(A, N, st); break;
            default: launch_raster<-1>(A, N, st); break;
        }
        OFK_LAUNCHED();
    }

    // ---- irregular part: boundary sites -> bins -> hull filter -> per-pixel point location
    IrrArgs I{flow, point_mask, d_folded, d_bins, d_coarse, d_sites, flow_sign, H, W, L.nbx, L.nby, L.ncx, L.ncy,
              (point_mask == nullptr && H >= 3 && W >= 3) ? 1 : 0};
    const long long cand = I.perimeter_only ? 2ll * W + 2ll * (H - 2) : (long long)H * W;
    dim3 sgrid((unsigned)((cand + 255) / 256), N);
    irr_sites_kernel<0><<<sgrid, 256, 0, st>>>(I);
    OFK_LAUNCHED();
    irr_scan_kernel<<<N, 1024, 0, st>>>(d_bins, L.nb);
    OFK_LAUNCHED();
    irr_sites_kernel<1><<<sgrid, 256, 0, st>>>(I);
    OFK_LAUNCHED();

    HullArgs Hh;
    Hh.flow = flow;
    Hh.folded = d_folded;
    Hh.bins = d_bins;
    Hh.sites = d_sites;
    Hh.ws = d_hullws;
    Hh.info = d_info;
    for (int k = 0; k < fwd::HULL_DIRS; ++k) {
        Hh.dirs.dx[k] = cos(2.0 * M_PI * k / fwd::HULL_DIRS);
        Hh.dirs.dy[k] = sin(2.0 * M_PI * k / fwd::HULL_DIRS);
    }
    Hh.sign = flow_sign;
    Hh.H = H;
    Hh.W = W;
    Hh.nb = L.nb;
    dim3 hgrid((unsigned)((cand + 256 * HULL_ITEMS - 1) / (256 * HULL_ITEMS)), N);
    hull_sites_kernel<0><<<hgrid, 256, 0, st>>>(Hh);
    OFK_LAUNCHED();
    hull_sites_kernel<1><<<hgrid, 256, 0, st>>>(Hh);
    OFK_LAUNCHED();
    hull_frame_kernel<0><<<N, 32, 0, st>>>(Hh);
    OFK_LAUNCHED();
    hull_sites_kernel<2><<<hgrid, 256, 0, st>>>(Hh);
    OFK_LAUNCHED();
    hull_frame_kernel<1><<<N, 32, 0, st>>>(Hh);
    OFK_LAUNCHED();

    SolveArgs S{payload, flow, payload_mask, d_folded, d_bins, d_coarse, d_sites, d_info, out, out_mask, d_cover,
                flow_sign, C, strict, H, W, L.nbx, L.nby, L.ncx, L.ncy};
    dim3 vgrid((W + 1023) / 1024, H, N);
    irr_solve_kernel<<<vgrid, 256, 0, st>>>(S);
    OFK_LAUNCHED();

    // ---- folding frames: redone by the order-independent resolve (no-ops for all other frames)
    if (H > 1 && W > 1) {
        unsigned int* winner = reinterpret_cast<unsigned int*>(d_sites);
        dim3 cgrid((unsigned)(((size_t)H * W + 1023) / 1024), N);
        legacy::fwd_clear<<<cgrid, 256, 0, st>>>(winner, (size_t)H * W, d_folded);
        OFK_LAUNCHED();
        dim3 grid((W - 1 + 31) / 32, (H - 1 + 7) / 8, N);
        legacy::fwd_scatter<<<grid, 256, 0, st>>>(flow, flow_sign, point_mask, winner, H, W, d_folded);
        OFK_LAUNCHED();
        dim3 ggrid((W + 31) / 32, (H + 7) / 8, N);
        const unsigned long long inv_w = ~0ull / (unsigned long long)W + 1ull;   // ceil(2^64 / W), W >= 2
        legacy::fwd_gather<<<ggrid, 256, 0, st>>>(payload, C, flow, flow_sign, payload_mask, winner, out, out_mask,
                                                 mask_rule, H, W, inv_w, d_folded);
        OFK_LAUNCHED();
    }
    return OFK_OK;
}

// ------------------------------------------------------------------------------------------- scattered -> scattered
// Barycentric interpolation of values attached to the displaced grid at arbitrary query points: replaces the direct
// `griddata(grid - A, A||mask, grid - B, 'linear', fill_value=0)` of combine_with mode 2 / ref 't'
// (flow_class.py:1398-1410) and the griddata calls of track_pts (utils.py:603,614). Same mesh as above (cells split
// along their Delaunay diagonal); the containing triangle of a query is found by a fixed-point walk
// p <- q - sign*flow(p) towards the source cell followed by an exact containment search around it.
namespace ofk {
namespace legacy {

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

}  // namespace legacy
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
    ofk::legacy::mesh_sample_kernel<<<grid, 128, 0, ofk::as_stream(stream)>>>(mesh_flow, mesh_sign, pos_f32, payload, C,
                                                                     payload_mask, query_flow, query_sign, query_pts,
                                                                     Q, out, out_maskval, found, H, W);
    OFK_LAUNCHED();
    return OFK_OK;
}
