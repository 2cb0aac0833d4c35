// pvr_slic.h -- SLICO superpixels of one 2D slice (zero-parameter SLIC, Achanta et al.), as the reference's PVR
// superpixel mode runs them per slice (source/reconstructionGPU2/runStackSLIC.cpp:666-842: grey -> RGB -> CIELAB,
// grid seeds, 10 iterations of SLICO with per-cluster adaptive compactness, connectivity enforcement).
#pragma once
#include <vector>

namespace svr {

// slice[y * X + x]; vmin / vmax = intensity range of the whole stack (the reference normalises to 0..255 with it);
// spx0 x spx1 = requested superpixel size.  Returns labels[y * X + x] (0 .. n_labels - 1) and sets *n_labels.
std::vector<int> slico_labels(const float* slice, int X, int Y, float vmin, float vmax, unsigned spx0, unsigned spx1, int* n_labels);

}  // namespace svr
