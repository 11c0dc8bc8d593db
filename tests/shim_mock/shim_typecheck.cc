// TEST (tests only): the -DB200SFM_WITH_GLOMAP branch of the shim against the API-faithful stub -- compiled with
// -fsyntax-only by tests/test_shim_cpu.py; instantiates the three estimators so every inline member is type-checked.
#include "estimators_shim.h"

bool Run(glomap::ViewGraph& vg, std::unordered_map<glomap::rig_t, glomap::Rig>& rigs,
         std::unordered_map<glomap::camera_t, glomap::Camera>& cameras,
         std::unordered_map<glomap::frame_t, glomap::Frame>& frames,
         std::unordered_map<glomap::image_t, glomap::Image>& images,
         std::unordered_map<glomap::track_t, glomap::Track>& tracks) {
  b200sfm_shim::RotationEstimatorOptions ro;
  b200sfm_shim::RotationEstimator ra(ro);
  b200sfm_shim::GlobalPositionerOptions go;
  b200sfm_shim::GlobalPositioner gp(go);
  b200sfm_shim::BundleAdjusterOptions bo;
  b200sfm_shim::BundleAdjuster ba(bo);
  ba.GetOptions().optimize_rotations = false;
  return ra.EstimateRotations(vg, rigs, frames, images) && gp.Solve(vg, rigs, cameras, frames, images, tracks) &&
         ba.Solve(rigs, cameras, frames, images, tracks);
}
