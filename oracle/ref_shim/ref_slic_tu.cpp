// ref_slic_tu.cpp -- TEST INFRASTRUCTURE.  Builds the reference's own SLICO code (source/reconstructionGPU2/runStackSLIC.cpp,
// compiled where it lies, unmodified) behind the inert IRTK stubs of irtk_stub/, and exposes the four algorithm members the
// per-slice driver calls (rgbtolab, getLABXYSeeds, PerformSuperpixelSLICO, EnforceSuperpixelConnectivity) through one C
// entry point.  The few driver lines between them (intensity normalisation, seed initialisation; runStackSLIC.cpp:735-782)
// need irtkGenericImage and are restated here.  Used by tests/golden/make_golden_slic.py and tests/test_slic_ref.py to pin
// host/pvr_slic.cc; never linked into the product.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>
using namespace std;          // the IRTK headers the reference includes inject it
#define REF_STR2(x) #x
#define REF_STR(x) REF_STR2(x)
#include REF_STR(REF_SLIC_CPP)

extern "C" int refslic_labels(const float* slice /* [Y][X], x fastest */, int X, int Y, float vmin, float vmax, unsigned spx0, unsigned spx1,
                              int* labels_out /* [Y][X] */)
{
    runStackSLIC<float> slic;
    const int width = Y, height = X, sz = X * Y;                     // the reference's orientation (runStackSLIC.cpp:703-705)
    const int numSuperpixels = (int)(sz / (spx0 * spx1));
    vector<int> rin(sz), gin(sz), bin(sz), klabels(sz), clabels(sz), seedIndices(sz);
    vector<double> lvec(sz), avec(sz), bvec(sz);
    int p = 0;
    for (int x = 0; x < X; ++x) for (int y = 0; y < Y; ++y) {
        const float v = slice[(size_t)y * X + x];
        rin[p] = gin[p] = bin[p] = (int) 255 * (v - vmin) / (vmax - vmin);
        p++;
    }
    slic.rgbtolab(rin.data(), gin.data(), bin.data(), sz, lvec.data(), avec.data(), bvec.data());
    const int step = sqrt((double)(sz) / (double)(numSuperpixels)) + 0.5;
    int numseeds = 0;
    slic.getLABXYSeeds(step, width, height, seedIndices.data(), &numseeds);
    vector<double> kx(numseeds), ky(numseeds), kl(numseeds), ka(numseeds), kb(numseeds);
    for (int k = 0; k < numseeds; k++) {
        kx[k] = seedIndices[k] % width; ky[k] = seedIndices[k] / width;
        kl[k] = lvec[seedIndices[k]]; ka[k] = avec[seedIndices[k]]; kb[k] = bvec[seedIndices[k]];
    }
    slic.PerformSuperpixelSLICO(lvec.data(), avec.data(), bvec.data(), kl.data(), ka.data(), kb.data(), kx.data(), ky.data(), width, height,
                                numseeds, klabels.data(), step);
    int n_final = 0;
    slic.EnforceSuperpixelConnectivity(klabels.data(), width, height, numSuperpixels, clabels.data(), &n_final);
    p = 0;
    for (int x = 0; x < X; ++x) for (int y = 0; y < Y; ++y) labels_out[(size_t)y * X + x] = clabels[p++];
    return n_final;
}
