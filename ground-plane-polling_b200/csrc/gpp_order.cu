// Scan order of the pair database (host code, once per database upload).
//
// The FAST / VERIFIED scans drop a whole row of 64 planes as soon as none of them can matter (gpp_poll3.cuh), so what a
// row costs depends on its WORST plane: in the order the databases ship in (random samples of KITTI road planes) 31 % of
// the rows of the all-six phase get past the first filter stage although only 2.5 % of the planes do.  Nothing in
// fit_road_planes.py:112-119 depends on the order the planes are visited in -- the arg-min is over the index, and both
// scans break ties by the ORIGINAL index explicitly -- so the pair database is stored in an order that makes rows
// homogeneous:
//   * a k-d tree over the normalised plane parameters (a, c, d) (b follows from the unit normal), each scaled by its
//     standard deviation: the node's rows are halved along the coordinate with the largest variance until a leaf is one
//     row of 64 planes;
//   * a stratified sample of the leaves (16 rows = 1024 planes, every (N/1024)-th plane of the tree order) goes FIRST: it
//     covers the parameter space evenly, so a scan has met a six-vote plane and a good residual bound after a few
//     rows, whatever the detection, and visits the homogeneous rows with that bound in hand.
// Measured on C4: rows of the all-six phase past stage 1 31 % -> 18 %, general-phase rows that leave
// after the bottom face 45 % -> 63 %; 4, 8, 16 and 32 sample rows time within 1 % of each other.  Databases below 32
// rows keep their order.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../include/gpp_debug.h"

namespace gpp {

namespace {

constexpr int kRow = 64;            // planes per row of the pair database
constexpr int kSeedRows = 16;
constexpr int kMinRows = 32;

struct Keys {
    std::vector<float> k[3];
};

void split(int *idx, int n_rows, int n, const Keys &K) {
    if (n_rows <= 1) {
        std::sort(idx, idx + n);
        return;
    }
    int dim = 0;
    double best = -1.0;
    for (int c = 0; c < 3; ++c) {
        double s = 0.0, s2 = 0.0;
        for (int i = 0; i < n; ++i) {
            const double v = K.k[c][idx[i]];
            s += v;
            s2 += v * v;
        }
        const double var = s2 / n - (s / n) * (s / n);
        if (var > best) {
            best = var;
            dim = c;
        }
    }
    const int left_rows = n_rows / 2, left = left_rows * kRow;       // the short row, if any, ends up last
    const std::vector<float> &key = K.k[dim];
    std::nth_element(idx, idx + left, idx + n, [&key](int a, int b) { return key[a] < key[b] || (key[a] == key[b] && a < b); });
    split(idx, left_rows, left, K);
    split(idx + left, n_rows - left_rows, n - left, K);
}

}  // namespace

// `rows`: N x 4 float32, row-major, as fed (not normalised).  order[position] = original plane index.
void scan_order(const float *rows, int n, std::vector<int32_t> &order) {
    order.resize((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    const int n_rows = (n + kRow - 1) / kRow;
    if (n_rows < kMinRows) return;
    if (const char *env = getenv("GPP_SCAN_ORDER"))      // measurements only: 0 keeps the index order
        if (atoi(env) == 0) return;
    Keys K;
    for (int c = 0; c < 3; ++c) K.k[c].resize((size_t)n);
    double mean[3] = {0, 0, 0}, m2[3] = {0, 0, 0};
    for (int j = 0; j < n; ++j) {
        // fit_road_planes.py:75-77 in double; only used to group similar planes, so a degenerate plane may sit anywhere
        const double a = rows[4 * j], b = rows[4 * j + 1], c = rows[4 * j + 2], d = rows[4 * j + 3];
        const double dir = b > 0 ? -1.0 : 1.0, rho = sqrt(a * a + b * b + c * c);
        double v[3] = {a * dir / rho, c * dir / rho, d * dir / rho};
        for (int i = 0; i < 3; ++i) {
            if (!(fabs(v[i]) < 1e30)) v[i] = 0.0;
            K.k[i][j] = (float)v[i];
            mean[i] += v[i];
            m2[i] += v[i] * v[i];
        }
    }
    for (int i = 0; i < 3; ++i) {
        const double mu = mean[i] / n, sd = sqrt(std::max(m2[i] / n - mu * mu, 0.0));
        const float scale = sd > 0 ? (float)(1.0 / sd) : 0.0f;
        for (int j = 0; j < n; ++j) K.k[i][j] = (float)((K.k[i][j] - mu) * scale);
    }
    std::vector<int> idx((size_t)n);
    std::iota(idx.begin(), idx.end(), 0);
    split(idx.data(), n_rows, n, K);
    // the stratified sample first, the rest in tree order
    int seed_rows = kSeedRows;
    if (const char *env = getenv("GPP_SEED_ROWS")) seed_rows = std::max(1, std::min(atoi(env), n_rows / 2));   // measurements only
    const int n_seed = seed_rows * kRow;
    std::vector<char> seed((size_t)n, 0);
    for (int i = 0; i < n_seed; ++i) seed[(size_t)(((double)i + 0.5) * n / n_seed)] = 1;
    size_t at = 0;
    for (int p = 0; p < n; ++p)
        if (seed[p]) order[at++] = idx[p];
    for (int p = 0; p < n; ++p)
        if (!seed[p]) order[at++] = idx[p];
}

}  // namespace gpp

extern "C" int gpp_debug_scan_order(const float *planes, int n_planes, int32_t *order) {
    if (!planes || !order || n_planes <= 0) return GPP_EINVAL;
    std::vector<int32_t> o;
    gpp::scan_order(planes, n_planes, o);
    std::copy(o.begin(), o.end(), order);
    return GPP_OK;
}
