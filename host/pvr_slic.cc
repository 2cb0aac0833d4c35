// pvr_slic.cc -- see pvr_slic.h.  The reference scans the slice with x as the slow index (runStackSLIC.cpp:735-747:
// "width" = Y, "height" = X, p = x * Y + y); the same orientation is kept here so that seeds, tie-breaks and the
// connectivity pass visit pixels in the same order.
#include "pvr_slic.h"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace svr {
namespace {

struct Lab { double l, a, b; };

// sRGB (D65) -> CIELAB of a grey value r = g = b = v (runStackSLIC.cpp:56-110)
Lab grey_to_lab(int v)
{
    const double c = v / 255.0;
    const double lin = c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4);
    const double X = lin * 0.4124564 + lin * 0.3575761 + lin * 0.1804375;
    const double Y = lin * 0.2126729 + lin * 0.7151522 + lin * 0.0721750;
    const double Z = lin * 0.0193339 + lin * 0.1191920 + lin * 0.9503041;
    const double eps = 0.008856, kappa = 903.3;
    auto f = [&](double t) { return t > eps ? std::pow(t, 1.0 / 3.0) : (kappa * t + 16.0) / 116.0; };
    const double fx = f(X / 0.950456), fy = f(Y / 1.0), fz = f(Z / 1.088754);
    return { 116.0 * fy - 16.0, 500.0 * (fx - fy), 200.0 * (fy - fz) };
}

}  // namespace

std::vector<int> slico_labels(const float* slice, int X, int Y, float vmin, float vmax, unsigned spx0, unsigned spx1, int* n_labels)
{
    const int width = Y, height = X, sz = X * Y;            // the reference's orientation: index = x * Y + y
    const int num_superpixels = std::max(1, (int)(sz / (spx0 * spx1)));
    std::vector<Lab> lab(sz);
    for (int x = 0, p = 0; x < X; ++x) for (int y = 0; y < Y; ++y, ++p) {
        const float v = slice[(size_t)y * X + x];
        lab[p] = grey_to_lab((int)(255 * (v - vmin) / (vmax - vmin)));     // float arithmetic, truncation (:741-743)
    }

    // ---- grid seeds (getLABXYSeeds, :112-153) ----------------------------------------------------------------------
    const int step = (int)(std::sqrt((double)sz / (double)num_superpixels) + 0.5);
    int xstrips = (int)(0.5 + (double)width / (double)step), ystrips = (int)(0.5 + (double)height / (double)step);
    int xerr = width - step * xstrips;
    if (xerr < 0) { xstrips--; xerr = width - step * xstrips; }
    int yerr = height - step * ystrips;
    if (yerr < 0) { ystrips--; yerr = height - step * ystrips; }
    xstrips = std::max(xstrips, 1); ystrips = std::max(ystrips, 1);
    const double xerrperstrip = (double)xerr / xstrips, yerrperstrip = (double)yerr / ystrips;
    struct Seed { double l, a, b, x, y; };
    std::vector<Seed> seeds;
    for (int y = 0; y < ystrips; ++y) for (int x = 0; x < xstrips; ++x) {
        const int sx = std::min(x * step + step / 2 + (int)(x * xerrperstrip), width - 1);
        const int sy = std::min(y * step + step / 2 + (int)(y * yerrperstrip), height - 1);
        const int i = sy * width + sx;
        seeds.push_back({ lab[i].l, lab[i].a, lab[i].b, (double)sx, (double)sy });
    }
    const int numk = (int)seeds.size();

    // ---- SLICO (PerformSuperpixelSLICO, :292-438): 10 iterations, D = d_lab / maxlab[k] + d_xy / step^2 ----------------
    std::vector<int> klabels(sz, -1);
    std::vector<double> distvec(sz), distlab(sz, DBL_MAX), maxlab(numk, 10.0 * 10.0);
    const double invxywt = 1.0 / ((double)step * step);
    for (int itr = 0; itr < 10; ++itr) {
        std::fill(distvec.begin(), distvec.end(), DBL_MAX);
        for (int n = 0; n < numk; ++n) {
            const Seed& s = seeds[n];
            const int x1 = std::max((int)(s.x - step), 0), y1 = std::max((int)(s.y - step), 0);
            const int x2 = std::min((int)(s.x + step), width), y2 = std::min((int)(s.y + step), height);
            for (int y = y1; y < y2; ++y) for (int x = x1; x < x2; ++x) {
                const int i = y * width + x;
                const double dl = lab[i].l - s.l, da = lab[i].a - s.a, db = lab[i].b - s.b;
                distlab[i] = dl * dl + da * da + db * db;
                const double distxy = (x - s.x) * (x - s.x) + (y - s.y) * (y - s.y);
                const double dist = distlab[i] / maxlab[n] + distxy * invxywt;
                if (dist < distvec[i]) { distvec[i] = dist; klabels[i] = n; }
            }
        }
        if (itr == 0) std::fill(maxlab.begin(), maxlab.end(), 1.0);
        for (int i = 0; i < sz; ++i) if (klabels[i] >= 0 && maxlab[klabels[i]] < distlab[i]) maxlab[klabels[i]] = distlab[i];
        std::vector<Seed> sum(numk, Seed{ 0, 0, 0, 0, 0 });
        std::vector<double> size(numk, 0.0);
        for (int r = 0, ind = 0; r < height; ++r) for (int c = 0; c < width; ++c, ++ind) {
            const int k = klabels[ind];
            if (k < 0) continue;
            sum[k].l += lab[ind].l; sum[k].a += lab[ind].a; sum[k].b += lab[ind].b; sum[k].x += c; sum[k].y += r;
            size[k] += 1.0;
        }
        for (int k = 0; k < numk; ++k) {
            const double inv = 1.0 / (size[k] <= 0 ? 1.0 : size[k]);
            seeds[k] = { sum[k].l * inv, sum[k].a * inv, sum[k].b * inv, sum[k].x * inv, sum[k].y * inv };
        }
    }

    // ---- EnforceSuperpixelConnectivity (:441-533): relabel 4-connected segments in scan order; a segment of at most
    // (sz / numSuperpixels) / 4 pixels takes the label of the last labelled neighbour seen ---------------------------------
    const int dx4[4] = { -1, 0, 1, 0 }, dy4[4] = { 0, -1, 0, 1 };
    const int supsz = sz / num_superpixels;
    std::vector<int> nlabels(sz, -1), xs, ys;
    int label = 0, adjlabel = 0;
    for (int j = 0, oindex = 0; j < height; ++j) for (int k = 0; k < width; ++k, ++oindex) {
        if (nlabels[oindex] >= 0) continue;
        nlabels[oindex] = label;
        xs.assign(1, k); ys.assign(1, j);
        for (int n = 0; n < 4; ++n) {
            const int x = k + dx4[n], y = j + dy4[n];
            if (x >= 0 && x < width && y >= 0 && y < height && nlabels[y * width + x] >= 0) adjlabel = nlabels[y * width + x];
        }
        for (size_t c = 0; c < xs.size(); ++c)
            for (int n = 0; n < 4; ++n) {
                const int x = xs[c] + dx4[n], y = ys[c] + dy4[n];
                if (x < 0 || x >= width || y < 0 || y >= height) continue;
                const int nindex = y * width + x;
                if (nlabels[nindex] < 0 && klabels[oindex] == klabels[nindex]) { xs.push_back(x); ys.push_back(y); nlabels[nindex] = label; }
            }
        if ((int)xs.size() <= (supsz >> 2)) {
            for (size_t c = 0; c < xs.size(); ++c) nlabels[ys[c] * width + xs[c]] = adjlabel;
            label--;
        }
        label++;
    }
    if (n_labels) *n_labels = label;

    std::vector<int> out(sz);
    for (int x = 0, p = 0; x < X; ++x) for (int y = 0; y < Y; ++y, ++p) out[(size_t)y * X + x] = nlabels[p];
    return out;
}

}  // namespace svr

// C entry point for the tests (tests/test_slic_ref.py compares it with the reference's own SLICO code)
extern "C" int svr_slico_labels(const float* slice, int X, int Y, float vmin, float vmax, unsigned spx0, unsigned spx1, int* labels_out)
{
    int n = 0;
    const std::vector<int> l = svr::slico_labels(slice, X, Y, vmin, vmax, spx0, spx1, &n);
    for (size_t i = 0; i < l.size(); ++i) labels_out[i] = l[i];
    return n;
}
