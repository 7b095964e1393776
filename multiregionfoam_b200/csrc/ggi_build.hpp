// ggi_build.hpp -- arithmetic and table construction of the GGI weight computation.
//
// The reference builds GGIInterpolation<standAlonePatch, standAlonePatch>(zoneA, zoneB, ..., SMALL, SMALL,
// rescale = true, BB_OCTREE) for every partitioned interface
// (src/numerics/interfaceToInterfaceMappings/ggiInterfaceToInterfaceMapping/ggiInterfaceToInterfaceMapping.C:62-77)
// and again whenever the interface moves (src/regionInterfaces/regionInterface/regionInterfaceType.C:483-511,
// 551-558).  GGIInterpolation itself is foam-extend 4.1 code (GGIInterpolationWeights.C, not in /root/reference);
// its published algorithm, restated here:
//   broad phase    candidate slave faces of a master face from bounding boxes (there: octree; here: uniform hash grid)
//   narrow phase   per pair: orthonormal basis (u, v, n) on the master face, both faces projected along n into (u, v),
//                  feature-angle test on the face normals, Sutherland-Hodgman clipping of the slave polygon by the
//                  (convex) master polygon, area of the intersection
//   weights        w = intersection area / master face area (both in the master plane), kept if above the
//                  non-overlap tolerance; rescaled to sum to one per face when asked
// The narrow phase and the per-face rescale run on the device (ggi_build.cuh); g++ compiles this same header into the
// CPU emulator of tests/cpp/ggi_emulate.cpp.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <unordered_map>
#include <vector>

#if defined(__CUDACC__)
#define B200_GHD __host__ __device__ __forceinline__
#else
#define B200_GHD inline
#endif

namespace ggib
{

constexpr int kMaxV = 8;              // points of one patch face (hexahedral / polyhedral patches: 3..8)
constexpr int kMaxClip = 2 * kMaxV;   // a convex m-gon clipped by m half-planes gains at most one point per half-plane
constexpr double kFeatureCos = 0.8;   // GGIInterpolation::featureCosTol_

struct V3
{
    double x, y, z;
};
B200_GHD V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
B200_GHD V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
B200_GHD V3 scale(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
B200_GHD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
B200_GHD V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
B200_GHD V3 point(const double* pts, int32_t i) { return {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]}; }

// face::centre / face::normal: triangle fan about the average point (a triangle directly)
B200_GHD void centre_normal(const double* pts, const int32_t* fp, int n, V3& ctr, V3& nrm)
{
    if (n == 3)
    {
        const V3 a = point(pts, fp[0]), b = point(pts, fp[1]), c = point(pts, fp[2]);
        ctr = scale(1.0 / 3.0, add(add(a, b), c));
        nrm = scale(0.5, cross(sub(b, a), sub(c, a)));
        return;
    }
    V3 avg = {0, 0, 0};
    for (int i = 0; i < n; i++) avg = add(avg, point(pts, fp[i]));
    avg = scale(1.0 / n, avg);
    V3 sumN = {0, 0, 0}, sumAc = {0, 0, 0};
    double sumA = 0;
    for (int i = 0; i < n; i++)
    {
        const V3 p = point(pts, fp[i]), q = point(pts, fp[(i + 1) % n]);
        const V3 c3 = add(add(p, q), avg);
        const V3 tn = cross(sub(q, p), sub(avg, p));
        const double a = sqrt(dot(tn, tn));
        sumN = add(sumN, tn);
        sumA += a;
        sumAc = add(sumAc, scale(a, c3));
    }
    ctr = sumA > 0 ? scale(1.0 / (3.0 * sumA), sumAc) : avg;
    nrm = scale(0.5, sumN);
}

B200_GHD double area2d(const double* x, const double* y, int n)
{
    double a = 0;
    for (int i = 0; i < n; i++)
    {
        const int j = i + 1 == n ? 0 : i + 1;
        a += x[i] * y[j] - x[j] * y[i];
    }
    return 0.5 * fabs(a);
}

// Intersection area of slave face (sfp, ns) with master face (mfp, nm) in the master plane; masterArea out.
// Returns 0 when the pair is rejected (feature angle) or the faces do not overlap.
B200_GHD double pair_area(const double* mPts, const int32_t* mfp, int nm, const double* sPts, const int32_t* sfp, int ns,
                          double& masterArea)
{
    V3 mc, mn, sc, sn;
    centre_normal(mPts, mfp, nm, mc, mn);
    centre_normal(sPts, sfp, ns, sc, sn);
    const double mmag = sqrt(dot(mn, mn)), smag = sqrt(dot(sn, sn));
    const V3 nhat = scale(1.0 / mmag, mn);
    // orthonormal basis on the master face: u along the first edge (made normal to n), v = n x u
    V3 e = sub(point(mPts, mfp[1]), point(mPts, mfp[0]));
    e = sub(e, scale(dot(e, nhat), nhat));
    const V3 uhat = scale(1.0 / sqrt(dot(e, e)), e);
    const V3 vhat = cross(nhat, uhat);
    double mx[kMaxV], my[kMaxV];
    for (int i = 0; i < nm; i++)
    {
        const V3 d = sub(point(mPts, mfp[i]), mc);
        mx[i] = dot(d, uhat);
        my[i] = dot(d, vhat);
    }
    masterArea = area2d(mx, my, nm);
    if (!(smag > 0) || fabs(dot(nhat, sn) / smag) < kFeatureCos) return 0.0;
    // subject polygon: the slave face projected along n into the master's (u, v)
    double ax[kMaxClip], ay[kMaxClip], bx[kMaxClip], by[kMaxClip];
    int na = ns;
    for (int i = 0; i < ns; i++)
    {
        const V3 d = sub(point(sPts, sfp[i]), mc);
        ax[i] = dot(d, uhat);
        ay[i] = dot(d, vhat);
    }
    // Sutherland-Hodgman: clip by every master edge (master polygon is counter-clockwise in (u, v) by construction)
    double *inx = ax, *iny = ay, *outx = bx, *outy = by;
    for (int k = 0; k < nm && na > 0; k++)
    {
        const int k1 = k + 1 == nm ? 0 : k + 1;
        const double ex = mx[k1] - mx[k], ey = my[k1] - my[k];
        int nb = 0;
        for (int i = 0; i < na; i++)
        {
            const int j = i + 1 == na ? 0 : i + 1;
            const double di = ex * (iny[i] - my[k]) - ey * (inx[i] - mx[k]); // >= 0: inside
            const double dj = ex * (iny[j] - my[k]) - ey * (inx[j] - mx[k]);
            if (di >= 0)
            {
                if (nb < kMaxClip)
                {
                    outx[nb] = inx[i];
                    outy[nb] = iny[i];
                    nb++;
                }
            }
            if ((di >= 0) != (dj >= 0))
            {
                const double t = di / (di - dj);
                if (nb < kMaxClip)
                {
                    outx[nb] = inx[i] + t * (inx[j] - inx[i]);
                    outy[nb] = iny[i] + t * (iny[j] - iny[i]);
                    nb++;
                }
            }
        }
        double* tx = inx;
        double* ty = iny;
        inx = outx;
        iny = outy;
        outx = tx;
        outy = ty;
        na = nb;
    }
    return na >= 3 ? area2d(inx, iny, na) : 0.0;
}

// Row i of the weight table: w[k] = area[k] / masterArea kept when above tol (else 0), rescaled to sum to one.
B200_GHD void row_weights(const double* area, double masterArea, int32_t begin, int32_t end, double tol, int rescale, double* w)
{
    double sum = 0;
    for (int32_t k = begin; k < end; k++)
    {
        const double wk = masterArea > 0 ? area[k] / masterArea : 0.0;
        w[k] = wk > tol ? wk : 0.0;
        sum += w[k];
    }
    if (rescale && sum > 0)
        for (int32_t k = begin; k < end; k++) w[k] = w[k] / sum;
}

// ---------------------------------------------------------------------------------------------- host: broad phase
struct Box
{
    double lo[3], hi[3];
};

inline Box face_box(const double* pts, const int32_t* fp, int n)
{
    Box b;
    for (int d = 0; d < 3; d++) b.lo[d] = b.hi[d] = pts[3 * fp[0] + d];
    for (int i = 1; i < n; i++)
        for (int d = 0; d < 3; d++)
        {
            b.lo[d] = std::min(b.lo[d], pts[3 * fp[i] + d]);
            b.hi[d] = std::max(b.hi[d], pts[3 * fp[i] + d]);
        }
    return b;
}

// Candidate slave faces (ascending) of every master face: bounding boxes, inflated by their own extent (the role of
// GGIInterpolation's bounding-box span factor), overlapping; uniform hash grid over the slave boxes.
inline void broad_phase(int32_t nM, const int32_t* mOff, const int32_t* mFp, const double* mPts, int32_t nS, const int32_t* sOff,
                        const int32_t* sFp, const double* sPts, std::vector<int32_t>& candOff, std::vector<int32_t>& cand)
{
    candOff.assign(nM + 1, 0);
    cand.clear();
    if (nM == 0 || nS == 0) return;
    auto inflated = [](Box b) {
        double ext = 0;
        for (int d = 0; d < 3; d++) ext = std::max(ext, b.hi[d] - b.lo[d]);
        for (int d = 0; d < 3; d++)
        {
            b.lo[d] -= 0.5 * ext;
            b.hi[d] += 0.5 * ext;
        }
        return b;
    };
    std::vector<Box> sb(nS);
    double h = 0, glo[3] = {1e300, 1e300, 1e300};
    for (int32_t j = 0; j < nS; j++)
    {
        sb[j] = inflated(face_box(sPts, sFp + sOff[j], sOff[j + 1] - sOff[j]));
        for (int d = 0; d < 3; d++)
        {
            h += (sb[j].hi[d] - sb[j].lo[d]) / (3.0 * nS);
            glo[d] = std::min(glo[d], sb[j].lo[d]);
        }
    }
    if (!(h > 0)) h = 1.0;
    auto cell = [&](double x, int d) { return (int64_t)std::floor((x - glo[d]) / h); };
    auto key = [](int64_t i, int64_t j, int64_t k) { return (uint64_t)((i * 73856093LL) ^ (j * 19349663LL) ^ (k * 83492791LL)); };
    std::unordered_map<uint64_t, std::vector<int32_t>> grid;
    grid.reserve((size_t)nS * 4);
    for (int32_t j = 0; j < nS; j++)
        for (int64_t a = cell(sb[j].lo[0], 0); a <= cell(sb[j].hi[0], 0); a++)
            for (int64_t b = cell(sb[j].lo[1], 1); b <= cell(sb[j].hi[1], 1); b++)
                for (int64_t c = cell(sb[j].lo[2], 2); c <= cell(sb[j].hi[2], 2); c++) grid[key(a, b, c)].push_back(j);
    std::vector<int32_t> row;
    for (int32_t i = 0; i < nM; i++)
    {
        const Box mb = inflated(face_box(mPts, mFp + mOff[i], mOff[i + 1] - mOff[i]));
        row.clear();
        for (int64_t a = cell(mb.lo[0], 0); a <= cell(mb.hi[0], 0); a++)
            for (int64_t b = cell(mb.lo[1], 1); b <= cell(mb.hi[1], 1); b++)
                for (int64_t c = cell(mb.lo[2], 2); c <= cell(mb.hi[2], 2); c++)
                {
                    auto it = grid.find(key(a, b, c));
                    if (it == grid.end()) continue;
                    for (int32_t j : it->second)
                    {
                        bool ov = true;
                        for (int d = 0; d < 3; d++) ov = ov && mb.lo[d] <= sb[j].hi[d] && sb[j].lo[d] <= mb.hi[d];
                        if (ov) row.push_back(j);
                    }
                }
        std::sort(row.begin(), row.end());
        row.erase(std::unique(row.begin(), row.end()), row.end());
        cand.insert(cand.end(), row.begin(), row.end());
        candOff[i + 1] = (int32_t)cand.size();
    }
}

// drop the zero-weight candidates: the GGI addressing / weights lists of the master faces
inline void compact(int32_t nM, const std::vector<int32_t>& candOff, const std::vector<int32_t>& cand, const double* w,
                    std::vector<int32_t>& off, std::vector<int32_t>& addr, std::vector<double>& weights)
{
    off.assign(nM + 1, 0);
    addr.clear();
    weights.clear();
    for (int32_t i = 0; i < nM; i++)
    {
        for (int32_t k = candOff[i]; k < candOff[i + 1]; k++)
            if (w[k] > 0)
            {
                addr.push_back(cand[k]);
                weights.push_back(w[k]);
            }
        off[i + 1] = (int32_t)addr.size();
    }
}

} // namespace ggib
