// TEST INFRASTRUCTURE -- not part of the product path.
//
// Thin pybind11 harness that exposes the *reference's own* per-stage CUDA launchers
// (compiled from /root/reference by oracle/build_ref.py into oracle/_ref/) so every
// intermediate of the 9-step pipeline can be dumped and compared stage by stage with
// oracle/stereo_oracle.c and with the sm_100a kernels.  Nothing here re-implements the
// algorithm: each function forwards to the reference launcher declared in
//   src/csrc/imageops/rgb_to_grayscale.hh:6-10, src/csrc/imageops/mean_pool.hh:6-10,
//   src/csrc/depth/kernels/ncc_matching_cost_volume_construction.hh:5-12,
//   src/csrc/depth/kernels/multi_block_matching_cost_aggregation.hh:5-13,
//   src/csrc/depth/kernels/wta_disparity_selection.hh:5-9,
//   src/csrc/depth/kernels/secondary_matching.hh:5-12,
//   src/csrc/depth/kernels/upscale_disparity_vertical_fill.hh:5-11,
//   src/csrc/depth/kernels/horizontal_disparity_fill.hh:5-10.
// The include paths below are resolved against the temporary patched copy of the
// reference tree that build_ref.py creates (never against files in this repo).
#include <torch/extension.h>

#include "imageops/rgb_to_grayscale.hh"
#include "imageops/mean_pool.hh"
#include "depth/kernels/ncc_matching_cost_volume_construction.hh"
#include "depth/kernels/multi_block_matching_cost_aggregation.hh"
#include "depth/kernels/wta_disparity_selection.hh"
#include "depth/kernels/secondary_matching.hh"
#include "depth/kernels/upscale_disparity_vertical_fill.hh"
#include "depth/kernels/horizontal_disparity_fill.hh"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("rgb_to_grayscale_inplace", &image_ops::rgb_to_grayscale_inplace);
    m.def("mean_pool_inplace", &image_ops::mean_pool_inplace);
    m.def("cost_volume", &ncc_matching_cost_volume_construction_cuda);
    m.def("aggregate", &multi_block_matching_cost_aggregation_cuda);
    m.def("wta", &wta_disparity_selection_cuda);
    m.def("secondary", &secondary_matching_cuda);
    m.def("upscale_vfill", &upscale_disparity_vertical_fill_cuda);
    m.def("hfill", &horizontal_disparity_fill_cuda);
}
