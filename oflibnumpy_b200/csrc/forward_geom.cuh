// Geometry of the source-referenced (forward) resampler, shared by the CUDA kernels (forward_s.cu) and by the host
// build of the same algorithm that the CPU test-suite compiles (tests/hostsim/forward_hostsim.cpp: test
// infrastructure, never loaded by the package).
//
// What is reproduced: scipy.interpolate.griddata(points = grid + flow, values, grid, 'linear') + nan_to_num at
// utils.py:237-258 of the reference, i.e. the Delaunay triangulation (Qhull) of the displaced pixel positions with
// barycentric interpolation inside each triangle and 0 outside the convex hull.
//
//   * Regular part. On a non-folding field the Delaunay triangulation contains the displaced pixel grid, every cell
//     split along its Delaunay diagonal. A cell whose four corners are all present ("intact") is rasterised directly:
//     raster_cell() below. Coverage follows a fill rule (edge functions evaluated with the endpoints in index order,
//     ties given to the triangle that traverses the edge in that order), so that every pixel inside the mesh is
//     produced by exactly ONE triangle: results can be written straight to the output, no id plane, no atomics.
//   * Irregular part (forward_irregular.cuh). Pixels no intact cell produces lie in a hole left by removed points
//     (`consider_mask`), in a pocket between the displaced frame border and its convex hull, or outside the hull.
//     Their Delaunay triangle only has "boundary sites" as vertices; it is located per pixel by a walk through the
//     Delaunay triangulation of those sites, found edge by edge with the empty-circle criterion.
//
// All predicates are float64 with a fixed operation sequence (explicit mul / fma), so the host build and the device
// build take identical decisions.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OFK_HD __host__ __device__ __forceinline__
#else
#define OFK_HD inline
#endif

namespace ofk {
namespace fwd {

struct P2 {
    double x, y;
};

// fixed-sequence float64 arithmetic: a*b, a*b+c with one rounding each, identical on host and device
OFK_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
OFK_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
OFK_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
OFK_HD double drcp(double a) {   // correctly rounded reciprocal on both sides
#if defined(__CUDA_ARCH__)
    return __drcp_rn(a);
#else
    return 1.0 / a;
#endif
}
OFK_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// position of pixel (row, col) displaced by sign * flow: int + float32 -> float64 exactly as numpy promotes
// `positions + flow_flat` (utils.py:242)
OFK_HD P2 displaced(float fx, float fy, int row, int col, float sign) {
    P2 p;
    p.x = dadd(static_cast<double>(col), static_cast<double>(sign * fx));
    p.y = dadd(static_cast<double>(row), static_cast<double>(sign * fy));
    return p;
}

// (b - a) x (c - a): > 0 when a, b, c are in the orientation of the undisplaced cell corners (a, b, d)
OFK_HD double orient(const P2& a, const P2& b, const P2& c) {
    const double bx = dsub(b.x, a.x), by = dsub(b.y, a.y), cx = dsub(c.x, a.x), cy = dsub(c.y, a.y);
    return dfma(bx, cy, -dmul(by, cx));
}

// > 0 iff d lies strictly inside the circumcircle of the positively oriented triangle (a, b, c)
OFK_HD double incircle(const P2& a, const P2& b, const P2& c, const P2& d) {
    const double ax = dsub(a.x, d.x), ay = dsub(a.y, d.y), bx = dsub(b.x, d.x), by = dsub(b.y, d.y),
                 cx = dsub(c.x, d.x), cy = dsub(c.y, d.y);
    const double a2 = dfma(ax, ax, dmul(ay, ay)), b2 = dfma(bx, bx, dmul(by, by)), c2 = dfma(cx, cx, dmul(cy, cy));
    const double m0 = dfma(by, c2, -dmul(b2, cy));
    const double m1 = dfma(bx, c2, -dmul(b2, cx));
    const double m2 = dfma(bx, cy, -dmul(by, cx));
    return dfma(a2, m2, dfma(ax, m0, -dmul(ay, m1)));
}

// In-circle test with a consistent answer for co-circular points: +1 = d inside the circumcircle of the positively
// oriented (a, b, c), -1 = outside. A determinant that vanishes to within its rounding is decided by a symbolic
// perturbation (every point lifted by an infinitesimal that shrinks rapidly with its id: the point with the smallest
// id decides, through its cofactor). The perturbed points are in general position, so every search sees the SAME
// triangulation of a co-circular set -- what keeps the point-location walk from cycling on exact lattices (similarity
// transforms of the pixel grid, float32 positions).
OFK_HD int incircle_sign(const P2& a, const P2& b, const P2& c, const P2& d, uint32_t ia, uint32_t ib, uint32_t ic_,
                         uint32_t id) {
    const double ax = dsub(a.x, d.x), ay = dsub(a.y, d.y), bx = dsub(b.x, d.x), by = dsub(b.y, d.y),
                 cx = dsub(c.x, d.x), cy = dsub(c.y, d.y);
    const double a2 = dfma(ax, ax, dmul(ay, ay)), b2 = dfma(bx, bx, dmul(by, by)), c2 = dfma(cx, cx, dmul(cy, cy));
    const double m0 = dfma(by, c2, -dmul(b2, cy));
    const double m1 = dfma(bx, c2, -dmul(b2, cx));
    const double m2 = dfma(bx, cy, -dmul(by, cx));
    const double det = dfma(a2, m2, dfma(ax, m0, -dmul(ay, m1)));
    // rounding of the expansion: a few ulps of the largest term
    const double mag = fabs(a2 * m2) + fabs(ax * m0) + fabs(ay * m1);
    if (fabs(det) > 4e-15 * mag) return det > 0 ? 1 : -1;
    // cofactors of the lifted coordinate: +orient(b,c,d), -orient(a,c,d), +orient(a,b,d), -orient(a,b,c)
    uint32_t ids[4] = {ia, ib, ic_, id};
    double cof[4] = {orient(b, c, d), -orient(a, c, d), orient(a, b, d), -orient(a, b, c)};
    for (int round = 0; round < 4; ++round) {
        int m = 0;
        for (int k = 1; k < 4; ++k)
            if (ids[k] < ids[m]) m = k;
        if (cof[m] != 0.0) return cof[m] > 0 ? 1 : -1;
        ids[m] = 0xffffffffu;
    }
    return -1;
}

// Edge function of point q against the edge u -> v, u being the endpoint with the SMALLER site index: both triangles
// sharing an edge evaluate the very same expression; the one traversing the edge as u -> v owns F == 0.
OFK_HD double edge_canon(const P2& u, const P2& v, double qx, double qy) {
    const double dx = dsub(v.x, u.x), dy = dsub(v.y, u.y), rx = dsub(qx, u.x), ry = dsub(qy, u.y);
    return dfma(dx, ry, -dmul(dy, rx));
}

// Corner naming of cell (i, j): a = (i, j), b = (i, j+1), c = (i+1, j), d = (i+1, j+1); site indices a < b < c < d.
// diag 0 splits along a-d: triangles (a, b, d) and (a, d, c); diag 1 along b-c: (a, b, c) and (b, d, c).
// Returns the diagonal (0 / 1) of the Delaunay split, or -1 when the displaced cell is not a positively oriented
// quadrilateral split that way (folding field).
// flip_tol (tests only, 0 in production): cells whose in-circle determinant is within +-flip_tol take the OTHER
// diagonal -- the alternative answer of a co-circular cell, against which the reference is compared where Qhull's pick
// is arbitrary.
OFK_HD int cell_diagonal(const P2& a, const P2& b, const P2& c, const P2& d, double (&area2)[2],
                         double flip_tol = 0.0) {
    const double o0a = orient(a, b, d), o0b = orient(a, d, c);   // diag 0 halves
    const double o1a = orient(a, b, c), o1b = orient(b, d, c);   // diag 1 halves
    const bool ok0 = o0a > 0 && o0b > 0, ok1 = o1a > 0 && o1b > 0;
    int diag;
    if (ok0 && ok1) {  // convex: Delaunay criterion, c inside the circumcircle of (a, b, d) -> flip to b-c
        const double ic = incircle(a, b, d, c);
        diag = ic > 0 ? 1 : 0;
        if (fabs(ic) <= flip_tol) diag ^= 1;
    } else if (ok0) {
        diag = 0;
    } else if (ok1) {
        diag = 1;
    } else {
        return -1;
    }
    area2[0] = diag ? o1a : o0a;
    area2[1] = diag ? o1b : o0b;
    return diag;
}

// Which triangle of the cell (split along `diag`) owns the point q: 0 / 1, or -1 when q belongs to another cell. The
// fill rule: an edge traversed from its smaller-index endpoint to the larger one includes F == 0, the reverse
// traversal excludes it, so a point on a shared edge or vertex has exactly one owner in the whole mesh.
OFK_HD int owner_triangle(const P2& a, const P2& b, const P2& c, const P2& d, int diag, double qx, double qy) {
    const double fab = edge_canon(a, b, qx, qy);
    const double fbd = edge_canon(b, d, qx, qy);
    const double fcd = edge_canon(c, d, qx, qy);
    const double fac = edge_canon(a, c, qx, qy);
    if (diag == 0) {
        const double fad = edge_canon(a, d, qx, qy);
        if (fab >= 0 && fbd >= 0 && fad < 0) return 0;     // (a, b, d): a->b, b->d canonical, d->a reversed
        if (fad >= 0 && fcd < 0 && fac < 0) return 1;      // (a, d, c): a->d canonical, d->c, c->a reversed
    } else {
        const double fbc = edge_canon(b, c, qx, qy);
        if (fab >= 0 && fbc >= 0 && fac < 0) return 0;     // (a, b, c): a->b, b->c canonical, c->a reversed
        if (fbd >= 0 && fcd < 0 && fbc < 0) return 1;      // (b, d, c): b->d canonical, d->c, c->b reversed
    }
    return -1;
}

// Barycentric weights of q in triangle `tri` of the cell: corner codes (0..3 = a..d) k[3] and weights w[3].
// area2 = twice the signed area of that triangle as cell_diagonal() returns it.
OFK_HD void triangle_weights(const P2& a, const P2& b, const P2& c, const P2& d, int diag, int tri, double area2,
                             double qx, double qy, int (&k)[3], double (&w)[3]) {
    double e0, e1;
    if (diag == 0) {
        if (tri == 0) {           // (a, b, d)
            e0 = edge_canon(b, d, qx, qy);
            e1 = -edge_canon(a, d, qx, qy);
            k[0] = 0; k[1] = 1; k[2] = 3;
        } else {                  // (a, d, c)
            e0 = -edge_canon(c, d, qx, qy);
            e1 = -edge_canon(a, c, qx, qy);
            k[0] = 0; k[1] = 3; k[2] = 2;
        }
    } else {
        if (tri == 0) {           // (a, b, c)
            e0 = edge_canon(b, c, qx, qy);
            e1 = -edge_canon(a, c, qx, qy);
            k[0] = 0; k[1] = 1; k[2] = 2;
        } else {                  // (b, d, c)
            e0 = -edge_canon(c, d, qx, qy);
            e1 = -edge_canon(b, c, qx, qy);
            k[0] = 1; k[1] = 3; k[2] = 2;
        }
    }
    const double r = drcp(area2);
    w[0] = dmul(e0, r);
    w[1] = dmul(e1, r);
    w[2] = dsub(dsub(1.0, w[0]), w[1]);
}

// Rasterises an intact cell: calls emit(x, y, k0, k1, k2, w0, w1, w2) once for every pixel the cell's two triangles
// own (k = corner codes 0..3 = a..d, w = barycentric weights). W, H: frame size (candidates are clipped to it).
template <class Emit>
OFK_HD void raster_cell(const P2& a, const P2& b, const P2& c, const P2& d, int diag, const double (&area2)[2],
                        int W, int H, Emit& emit) {
    const double xlo = fmin(fmin(a.x, b.x), fmin(c.x, d.x)), xhi = fmax(fmax(a.x, b.x), fmax(c.x, d.x));
    const double ylo = fmin(fmin(a.y, b.y), fmin(c.y, d.y)), yhi = fmax(fmax(a.y, b.y), fmax(c.y, d.y));
    // comparisons in floating point first: the positions are unbounded, the integer casts must not overflow
    if (!(xhi >= 0.0 && yhi >= 0.0 && xlo <= (double)(W - 1) && ylo <= (double)(H - 1))) return;
    const int x0 = (int)ceil(fmax(xlo, 0.0)), x1 = (int)floor(fmin(xhi, (double)(W - 1)));
    const int y0 = (int)ceil(fmax(ylo, 0.0)), y1 = (int)floor(fmin(yhi, (double)(H - 1)));
    for (int y = y0; y <= y1; ++y) {
        for (int x = x0; x <= x1; ++x) {
            const int tri = owner_triangle(a, b, c, d, diag, (double)x, (double)y);
            if (tri < 0) continue;
            int k[3];
            double w[3];
            triangle_weights(a, b, c, d, diag, tri, area2[tri], (double)x, (double)y, k, w);
            emit(x, y, k[0], k[1], k[2], w[0], w[1], w[2]);
        }
    }
}

// The reference's `== 1` on the resampled mask channel (flow_class.py:668): float payloads are cast to float32
// first (utils.py:258), integer payloads are rounded (np.round(m) == 1 <=> m > 0.5). m = sum of the weights of the
// vertices whose payload mask is set.
OFK_HD bool mask_rule_s(double m, int rule_strict) {
    return rule_strict ? (static_cast<float>(m) == 1.0f) : (m > 0.5);
}

// One output pixel: payload channels interpolated in float64 and cast to float32 (griddata works in float64, the
// reference casts back at utils.py:258), validity from the weights of the vertices whose payload mask is set.
OFK_HD void interp_store(const float* p0, const float* p1, const float* p2, bool m0, bool m1, bool m2, double w0,
                         double w1, double w2, int C, float* out_px, uint8_t* out_mask_px, int rule_strict) {
    for (int c = 0; c < C; ++c) {
        const double v = dfma(w2, (double)p2[c], dfma(w1, (double)p1[c], dmul(w0, (double)p0[c])));
        out_px[c] = (float)v;
    }
    if (out_mask_px != nullptr) {
        bool ok = true;
        if (!(m0 && m1 && m2)) {
            const double m = dadd(dadd(m0 ? w0 : 0.0, m1 ? w1 : 0.0), m2 ? w2 : 0.0);
            ok = mask_rule_s(m, rule_strict);
        }
        *out_mask_px = ok ? 1 : 0;
    }
}

}  // namespace fwd
}  // namespace ofk
