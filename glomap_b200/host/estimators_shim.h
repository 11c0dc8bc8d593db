// estimators_shim.h -- C++ drop-in for glomap's three estimator classes over
// the C ABI (include/b200sfm.h).  Same class names, constructor / method
// signatures, option fields and bool-return error behaviour as
//   glomap::RotationEstimator   glomap/estimators/global_rotation_averaging.h:39-87
//   glomap::GlobalPositioner    glomap/estimators/global_positioning.h:9-70
//   glomap::BundleAdjuster      glomap/estimators/bundle_adjustment.h:12-51
// Each Solve flattens the unordered_map world into SoA in SORTED-ID order
// (deterministic, unlike the reference's hash-map order), calls the GPU solver
// and scatters the results back in place.  Known (constant) camera rigs are
// supported in BundleAdjuster and GlobalPositioner (frames = pose blocks, every
// image carries its sensor's cam_from_rig); optimize_rig_poses (BA) marks the non-reference sensors as unknowns
// and writes the optimised cam_from_rig back.  Rigs with a not-yet-calibrated sensor (RA / GP: unknown cam_from_rig
// blocks, available at the C ABI and in the Python host) return false with a message on stderr.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <array>
#include <map>
#include <limits>
#include <queue>
#include <set>
#include <random>
#include <string>
#include <vector>

#include "../../include/b200sfm.h"
#include "scene_min.h"

namespace b200sfm_shim {
using namespace b200host;

// One context per CUDA device, created on first use.  `gpu_index` is the reference's option string
// (bundle_adjustment.h:24, global_positioning.h:42: "-1" = the default device, otherwise the first index of a
// comma-separated list -- the reference hands the list to Ceres/cuDSS, which uses one device as well).
inline b200sfm_ctx* DefaultContext(const std::string& gpu_index = "-1") {
  static std::map<int, b200sfm_ctx*> ctxs;
  int device = 0;
  try {
    device = std::stoi(gpu_index);
  } catch (...) {
    device = -1;
  }
  if (device < 0) device = 0;
  auto it = ctxs.find(device);
  if (it != ctxs.end()) return it->second;
  b200sfm_ctx* ctx = nullptr;
  if (b200sfm_create(device, &ctx) != B200SFM_OK) {
    std::fprintf(stderr, "b200sfm: no CUDA device %d / context creation failed (there is no CPU fallback)\n", device);
    return nullptr;   // not cached: a later call may name a valid device
  }
  ctxs[device] = ctx;
  return ctx;
}

// small quaternion helpers (xyzw, Eigen coeffs() order)
inline void QuatMul(const double* a, const double* b, double* o) {   // o = a (x) b
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  const double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  const double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
inline void QuatConj(const double* a, double* o) { o[0] = -a[0]; o[1] = -a[1]; o[2] = -a[2]; o[3] = a[3]; }
// colmap::AverageQuaternions with unit weights: dominant eigenvector of sum q q^T (sign-invariant), by power iteration
inline void AverageQuaternions(const std::vector<std::array<double, 4>>& qs, double* out) {
  double M[4][4] = {};
  for (const auto& q : qs)
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) M[r][c] += q[r] * q[c];
  double v[4] = {qs[0][0], qs[0][1], qs[0][2], qs[0][3]};
  for (int it = 0; it < 200; ++it) {
    double u[4] = {0, 0, 0, 0};
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) u[r] += M[r][c] * v[c];
    const double n = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + u[3] * u[3]);
    if (!(n > 0)) break;
    for (int k = 0; k < 4; ++k) v[k] = u[k] / n;
  }
  for (int k = 0; k < 4; ++k) out[k] = v[k];
}

inline void QuatToR(const double* q, double R[9]) {
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double x = q[0] / n, y = q[1] / n, z = q[2] / n, w = q[3] / n;
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}

// ---------------------------------------------------------------------------
struct OptimizationBaseOptions {                 // optimization_base.h:10-24
  double thres_loss_function = 1e-1;
  struct SolverOptions {
    int max_num_iterations = 100;
    double function_tolerance = 1e-5;
    double gradient_tolerance = 1e-10;
    double parameter_tolerance = 1e-8;
  } solver_options;
};

struct BundleAdjusterOptions : public OptimizationBaseOptions {   // bundle_adjustment.h:12-37
  bool optimize_rig_poses = false;
  bool optimize_rotations = true;
  bool optimize_translation = true;
  bool optimize_intrinsics = true;
  bool optimize_principal_point = false;
  bool optimize_points = true;
  bool use_gpu = true;
  std::string gpu_index = "-1";
  int min_num_images_gpu_solver = 50;
  int min_num_view_per_track = 3;
  // PCG knobs of this implementation
  double pcg_rel_tolerance = 1e-2;
  int pcg_max_iterations = 500;
  BundleAdjusterOptions() {
    thres_loss_function = 1.;
    solver_options.max_num_iterations = 200;
  }
};

class BundleAdjuster {
 public:
  BundleAdjuster(const BundleAdjusterOptions& options) : options_(options) {}
  BundleAdjusterOptions& GetOptions() { return options_; }
  b200sfm_lm_stats summary{};

  // bundle_adjustment.cc:11-106
  bool Solve(std::unordered_map<rig_t, Rig>& rigs, std::unordered_map<camera_t, Camera>& cameras,
             std::unordered_map<frame_t, Frame>& frames, std::unordered_map<image_t, Image>& images,
             std::unordered_map<track_t, Track>& tracks) {
    if (images.empty()) { std::fprintf(stderr, "Number of images = 0\n"); return false; }     // .cc:17-20
    if (tracks.empty()) { std::fprintf(stderr, "Number of tracks = 0\n"); return false; }     // .cc:21-24
    b200sfm_ctx* ctx = DefaultContext(options_.gpu_index);
    if (!ctx) return false;
    // frames / cameras / tracks in sorted-id order
    std::map<frame_t, Frame*> fsorted;
    for (auto& [id, f] : frames) fsorted[id] = &f;
    std::map<frame_t, int> fidx;
    for (auto& [id, f] : fsorted) { const int i = (int)fidx.size(); fidx[id] = i; }
    std::map<camera_t, Camera*> csorted;
    for (auto& [id, c] : cameras) csorted[id] = &c;
    std::map<camera_t, int> cidx;
    for (auto& [id, c] : csorted) { const int i = (int)cidx.size(); cidx[id] = i; }
    const int C = (int)fsorted.size(), K = (int)csorted.size();
    std::vector<double> quat(4 * (size_t)C), trans(3 * (size_t)C), intr((size_t)K * B200SFM_INTR_STRIDE, 0.0);
    std::vector<int32_t> cam_intr(C, 0), intr_model(K);
    std::vector<uint8_t> mask(C, 0);
    for (auto& [id, f] : fsorted) {
      const int i = fidx[id];
      for (int k = 0; k < 4; ++k) quat[4 * i + k] = f->RigFromWorld().rotation.coeffs().data()[k];
      for (int k = 0; k < 3; ++k) trans[3 * i + k] = f->RigFromWorld().translation[k];
    }
    // the gauge frame is chosen below, once it is known which frames carry observations (.cc:252-266)
    for (auto& [id, c] : csorted) {
      const int k = cidx[id];
      intr_model[k] = static_cast<int32_t>(c->model_id);   // colmap::CameraModelId is an enum class
      for (size_t j = 0; j < c->params.size() && j < B200SFM_INTR_STRIDE; ++j) intr[(size_t)k * B200SFM_INTR_STRIDE + j] = c->params[j];
    }
    // sensors = (rig, camera) pairs in sorted order; trivial frames use the identity cam_from_rig
    bool any_rig = false;
    for (auto& [id, im] : images) any_rig = any_rig || !im.HasTrivialFrame();
    std::map<std::pair<rig_t, camera_t>, int> sidx;
    std::vector<double> sensor_q, sensor_t;
    std::vector<int32_t> sensor_intr;
    std::vector<uint8_t> sensor_var;   // optimize_rig_poses: every non-reference sensor is an unknown (.cc:162-180,296-308)
    if (any_rig) {
      for (auto& [id, im] : images) sidx[{frames[im.frame_id].RigId(), im.camera_id}] = 0;
      int n = 0;
      for (auto& [key, idx] : sidx) {
        idx = n++;
        Rigid3d cfr;   // identity
        bool trivial = true;
        for (auto& [iid, im] : images)
          if (frames[im.frame_id].RigId() == key.first && im.camera_id == key.second) { trivial = im.HasTrivialFrame(); break; }
        if (!trivial) cfr = b200host_adapt::CamFromRig(rigs[key.first], key.second);
        for (int k = 0; k < 4; ++k) sensor_q.push_back(cfr.rotation.coeffs().data()[k]);
        for (int k = 0; k < 3; ++k) sensor_t.push_back(cfr.translation[k]);
        sensor_intr.push_back(cidx[key.second]);
        // NonRefSensors() of the rig (.cc:299); Image::HasTrivialFrame() is exactly IsRefSensor (scene/image.h:73-76)
        sensor_var.push_back((options_.optimize_rig_poses && !b200host_adapt::IsRefSensor(rigs[key.first], key.second)) ? 1 : 0);
      }
      if (n > 65535) { std::fprintf(stderr, "b200sfm: too many rig sensors\n"); return false; }
    } else {
      for (auto& [id, im] : images) cam_intr[fidx[im.frame_id]] = cidx[im.camera_id];
    }
    std::map<track_t, Track*> tsorted;
    for (auto& [id, t] : tracks) tsorted[id] = &t;
    const int P = (int)tsorted.size();
    std::vector<int64_t> ptb(1, 0);
    std::vector<int32_t> obs_cam;
    std::vector<uint16_t> obs_sensor;
    std::vector<double> obs_xy, points(3 * (size_t)P);
    int p = 0;
    for (auto& [id, t] : tsorted) {
      // .cc:122: the track is skipped on ITS observation count; observations of missing images are dropped afterwards
      // (.cc:125), so the device gets min_num_view_per_track = 1 and a skipped track simply carries no observation
      const bool keep = (int)t->observations.size() >= options_.min_num_view_per_track;
      for (const auto& ob : t->observations) {
        if (!keep) break;
        auto it = images.find(ob.first);
        if (it == images.end()) continue;                                                     // .cc:125
        obs_cam.push_back(fidx[it->second.frame_id]);
        if (any_rig) obs_sensor.push_back((uint16_t)sidx[{frames[it->second.frame_id].RigId(), it->second.camera_id}]);
        obs_xy.push_back(it->second.features[ob.second][0]);
        obs_xy.push_back(it->second.features[ob.second][1]);
      }
      ptb.push_back((int64_t)obs_cam.size());
      for (int k = 0; k < 3; ++k) points[3 * (size_t)p + k] = t->xyz[k];
      ++p;
    }
    {   // .cc:252-266: the first frame (map order; here sorted-id order) that HAS a parameter block is held constant
      std::vector<uint8_t> used(C, 0);
      for (int32_t f : obs_cam) used[f] = 1;
      for (int i = 0; i < C; ++i)
        if (used[i]) { mask[i] = 3; break; }
    }
    b200sfm_ba_opts o;
    b200sfm_ba_default_opts(&o);
    o.optimize_rig_poses = options_.optimize_rig_poses; o.optimize_rotations = options_.optimize_rotations;
    o.optimize_translation = options_.optimize_translation; o.optimize_intrinsics = options_.optimize_intrinsics;
    o.optimize_principal_point = options_.optimize_principal_point; o.optimize_points = options_.optimize_points;
    o.min_num_view_per_track = 1;   // the track-length rule was applied above, on track.observations.size()
    o.max_num_iterations = options_.solver_options.max_num_iterations;
    o.thres_loss_function = options_.thres_loss_function;
    o.function_tolerance = options_.solver_options.function_tolerance;
    o.gradient_tolerance = options_.solver_options.gradient_tolerance;
    o.parameter_tolerance = options_.solver_options.parameter_tolerance;
    o.pcg_rel_tolerance = options_.pcg_rel_tolerance; o.pcg_max_iterations = options_.pcg_max_iterations;
    int rc;
    if (any_rig) {   // known rigs: resident-problem path (bundle_adjustment.cc:147-161)
      b200sfm_ba_problem* prob = nullptr;
      rc = b200sfm_ba_problem_create_rig(ctx, C, P, (int64_t)obs_cam.size(), K, (int32_t)sensor_intr.size(), ptb.data(),
                                         obs_cam.data(), obs_sensor.data(), obs_xy.data(), sensor_q.data(), sensor_t.data(),
                                         sensor_intr.data(), intr_model.data(), mask.data(), o.min_num_view_per_track, &prob);
      if (rc == B200SFM_OK) rc = b200sfm_ba_problem_set_state(prob, intr.data(), quat.data(), trans.data(), points.data());
      if (rc == B200SFM_OK && options_.optimize_rig_poses) rc = b200sfm_ba_problem_set_sensor_variable(prob, sensor_var.data());
      if (rc == B200SFM_OK) rc = b200sfm_ba_problem_solve(prob, &o, &summary);
      if (rc == B200SFM_OK) rc = b200sfm_ba_problem_get_state(prob, intr.data(), quat.data(), trans.data(), points.data());
      if (rc == B200SFM_OK && options_.optimize_rig_poses) {   // the optimised cam_from_rig back into the rigs, in place
        rc = b200sfm_ba_problem_get_sensor_poses(prob, sensor_q.data(), sensor_t.data());
        if (rc == B200SFM_OK)
          for (auto& [key, idx] : sidx) {
            if (!sensor_var[idx]) continue;
            Rigid3d cfr;
            for (int k = 0; k < 4; ++k) cfr.rotation.coeffs().data()[k] = sensor_q[4 * (size_t)idx + k];
            for (int k = 0; k < 3; ++k) cfr.translation[k] = sensor_t[3 * (size_t)idx + k];
            b200host_adapt::SetCamFromRig(rigs[key.first], key.second, cfr);
          }
      }
      b200sfm_ba_problem_free(prob);
    } else {
      rc = b200sfm_ba_solve(ctx, &o, C, P, (int64_t)obs_cam.size(), K, ptb.data(), obs_cam.data(), obs_xy.data(),
                            cam_intr.data(), intr_model.data(), intr.data(), quat.data(), trans.data(), mask.data(),
                            points.data(), &summary);
    }
    if (rc != B200SFM_OK) { std::fprintf(stderr, "b200sfm_ba_solve: %s\n", b200sfm_last_error(ctx)); return false; }
    for (auto& [id, f] : fsorted) {                                                           // results in place (.cc:140-146)
      const int i = fidx[id];
      for (int k = 0; k < 4; ++k) f->RigFromWorld().rotation.coeffs().data()[k] = quat[4 * i + k];
      for (int k = 0; k < 3; ++k) f->RigFromWorld().translation[k] = trans[3 * i + k];
    }
    p = 0;
    for (auto& [id, t] : tsorted) { for (int k = 0; k < 3; ++k) t->xyz[k] = points[3 * (size_t)p + k]; ++p; }
    for (auto& [id, c] : csorted)
      for (size_t j = 0; j < c->params.size() && j < B200SFM_INTR_STRIDE; ++j) c->params[j] = intr[(size_t)cidx[id] * B200SFM_INTR_STRIDE + j];
    return summary.usable != 0;                                                               // .cc:105
  }

 private:
  BundleAdjusterOptions options_;
};

// ---------------------------------------------------------------------------
struct GlobalPositionerOptions : public OptimizationBaseOptions {   // global_positioning.h:9-54
  enum ConstraintType { ONLY_POINTS, ONLY_CAMERAS, POINTS_AND_CAMERAS_BALANCED, POINTS_AND_CAMERAS };
  bool generate_random_positions = true, generate_random_points = true, generate_scales = true;
  bool optimize_positions = true, optimize_points = true, optimize_scales = true;
  bool use_gpu = true;
  std::string gpu_index = "-1";
  int min_num_images_gpu_solver = 50;
  int min_num_view_per_track = 3;
  unsigned seed = 1;
  ConstraintType constraint_type = ONLY_POINTS;
  double constraint_reweight_scale = 1.0;
  double pcg_rel_tolerance = 1e-2;
  int pcg_max_iterations = 1000;
  GlobalPositionerOptions() { thres_loss_function = 1e-1; }
};

class GlobalPositioner {
 public:
  GlobalPositioner(const GlobalPositionerOptions& options) : options_(options) { random_generator_.seed(options_.seed); }
  GlobalPositionerOptions& GetOptions() { return options_; }
  b200sfm_lm_stats summary{};

  // global_positioning.cc:28-93 (ONLY_POINTS, trivial rigs)
  bool Solve(const ViewGraph& view_graph, std::unordered_map<rig_t, Rig>& rigs,
             std::unordered_map<camera_t, Camera>& cameras, std::unordered_map<frame_t, Frame>& frames,
             std::unordered_map<image_t, Image>& images, std::unordered_map<track_t, Track>& tracks) {
    (void)view_graph;
    if (images.empty()) { std::fprintf(stderr, "Number of images = 0\n"); return false; }     // .cc:37-40
    if (tracks.empty()) { std::fprintf(stderr, "Number of tracks = 0\n"); return false; }     // .cc:46-50
    if (options_.constraint_type != GlobalPositionerOptions::ONLY_POINTS) {
      std::fprintf(stderr, "b200sfm: only ONLY_POINTS is implemented\n");
      return false;
    }
    b200sfm_ctx* ctx = DefaultContext(options_.gpu_index);
    if (!ctx) return false;
    std::map<frame_t, Frame*> fsorted;
    for (auto& [id, f] : frames) fsorted[id] = &f;
    std::map<frame_t, int> fidx;
    for (auto& [id, f] : fsorted) { const int i = (int)fidx.size(); fidx[id] = i; }
    const int C = (int)fsorted.size();
    std::uniform_real_distribution<double> U(-1, 1);
    std::vector<double> centers(3 * (size_t)C), Rm(9 * (size_t)C);
    std::vector<uint8_t> calibrated(C, 1);
    for (auto& [id, f] : fsorted) {
      const int i = fidx[id];
      QuatToR(f->RigFromWorld().rotation.coeffs().data(), &Rm[9 * (size_t)i]);
      for (int k = 0; k < 3; ++k) {
        if (options_.generate_random_positions && options_.optimize_positions) {
          centers[3 * i + k] = 100.0 * U(random_generator_);                                  // .cc:158-159
        } else {                                                                              // CenterFromPose: -R^T t
          const double* R = &Rm[9 * (size_t)i];
          const auto& t = f->RigFromWorld().translation;
          centers[3 * i + k] = -(R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2]);
        }
      }
    }
    bool any_rig = false;
    for (auto& [id, im] : images) {
      calibrated[fidx[im.frame_id]] = cameras[im.camera_id].has_prior_focal_length ? 1 : 0;
      any_rig = any_rig || !im.HasTrivialFrame();
    }
    std::vector<double> obs_off;      // known rigs: R_cw^T t_cam_from_rig per observation (.cc:339-345)
    std::vector<uint8_t> obs_cal;     // the loss is chosen per CAMERA (.cc:313-316)
    // sensors whose cam_from_rig translation is still NaN (rotation averaging estimated their rotation only): their
    // centre in the rig frame is an unknown of this solve, RigUnknownBATAPairwiseDirectionError (.cc:355-372)
    std::map<std::pair<rig_t, camera_t>, int> usens;
    std::vector<int32_t> obs_usens;
    std::map<track_t, Track*> tsorted;
    for (auto& [id, t] : tracks) tsorted[id] = &t;
    const int P = (int)tsorted.size();
    std::vector<int64_t> ptb(1, 0);
    std::vector<int32_t> obs_cam;
    std::vector<double> obs_dir, points(3 * (size_t)P);
    int p = 0;
    for (auto& [id, t] : tsorted) {
      const bool keep = (int)t->observations.size() >= options_.min_num_view_per_track;        // .cc:257-258
      for (const auto& ob : t->observations) {
        if (!keep) break;
        auto it = images.find(ob.first);
        if (it == images.end() || !it->second.IsRegistered()) continue;                      // .cc:279-282
        const auto& b = it->second.features_undist[ob.second];
        if (std::isnan(b[0]) || std::isnan(b[1]) || std::isnan(b[2])) continue;               // .cc:286-292
        const int ci = fidx[it->second.frame_id];
        const double* R = &Rm[9 * (size_t)ci];
        if (any_rig) {
          // cam_from_world = cam_from_rig * rig_from_world;  t_obs = R_cw^T b,  t_rig = R_cw^T t_cam_from_rig
          Rigid3d cfr;
          if (!it->second.HasTrivialFrame())
            cfr = b200host_adapt::CamFromRig(rigs[frames[it->second.frame_id].RigId()], it->second.camera_id);
          int us = -1;
          if (std::isnan(cfr.translation[0]) || std::isnan(cfr.translation[1]) || std::isnan(cfr.translation[2])) {
            const auto key = std::make_pair(frames[it->second.frame_id].RigId(), it->second.camera_id);
            auto u = usens.find(key);
            if (u == usens.end()) u = usens.emplace(key, (int)usens.size()).first;
            us = u->second;
            cfr.translation[0] = cfr.translation[1] = cfr.translation[2] = 0.0;   // no known offset: the centre is the unknown
          }
          obs_usens.push_back(us);
          double Rs[9], bb[3], tt[3];
          QuatToR(cfr.rotation.coeffs().data(), Rs);
          for (int k = 0; k < 3; ++k) {   // R_cr^T b, R_cr^T t_cr
            bb[k] = Rs[k] * b[0] + Rs[3 + k] * b[1] + Rs[6 + k] * b[2];
            tt[k] = Rs[k] * cfr.translation[0] + Rs[3 + k] * cfr.translation[1] + Rs[6 + k] * cfr.translation[2];
          }
          for (int k = 0; k < 3; ++k) {
            obs_dir.push_back(R[k] * bb[0] + R[3 + k] * bb[1] + R[6 + k] * bb[2]);
            obs_off.push_back(R[k] * tt[0] + R[3 + k] * tt[1] + R[6 + k] * tt[2]);
          }
          obs_cal.push_back(cameras[it->second.camera_id].has_prior_focal_length ? 1 : 0);
        } else {
          for (int k = 0; k < 3; ++k) obs_dir.push_back(R[k] * b[0] + R[3 + k] * b[1] + R[6 + k] * b[2]);   // R^T b (.cc:294-296)
        }
        obs_cam.push_back(ci);
      }
      ptb.push_back((int64_t)obs_cam.size());
      const bool rnd = options_.optimize_points && options_.generate_random_points &&
                       (int)t->observations.size() >= options_.min_num_view_per_track;
      for (int k = 0; k < 3; ++k) points[3 * (size_t)p + k] = rnd ? 100.0 * U(random_generator_) : t->xyz[k];   // .cc:261-264
      if (rnd) t->is_initialized = true;
      ++p;
    }
    std::vector<double> scales(obs_cam.size(), 1.0);                                          // .cc:298
    b200sfm_gp_opts o;
    b200sfm_gp_default_opts(&o);
    o.optimize_positions = options_.optimize_positions; o.optimize_points = options_.optimize_points;
    o.optimize_scales = options_.optimize_scales;
    o.min_num_view_per_track = 1;   // the track-length rule was applied above, on track.observations.size()
    o.max_num_iterations = options_.solver_options.max_num_iterations;
    o.thres_loss_function = options_.thres_loss_function;
    o.function_tolerance = options_.solver_options.function_tolerance;
    o.pcg_rel_tolerance = options_.pcg_rel_tolerance; o.pcg_max_iterations = options_.pcg_max_iterations;
    int rc;
    if (any_rig) {   // RigBATA with the rig scales held constant (.cc:325-346,493-497)
      b200sfm_gp_problem* prob = nullptr;
      rc = b200sfm_gp_problem_create(ctx, C, P, (int64_t)obs_cam.size(), ptb.data(), obs_cam.data(), obs_dir.data(),
                                     calibrated.data(), nullptr, o.min_num_view_per_track, &prob);
      if (rc == B200SFM_OK) rc = b200sfm_gp_problem_set_rig_terms(prob, obs_off.data(), obs_cal.data());
      // unknown sensor centres: U(-1, 1)^3 when the positions are optimised (.cc:440-453), in order of first appearance
      std::vector<double> ucen(3 * usens.size(), 0.0);
      if (!usens.empty()) {
        if (options_.optimize_positions)
          for (double& v : ucen) v = U(random_generator_);
        if (rc == B200SFM_OK)
          rc = b200sfm_gp_problem_set_rig_unknown(prob, (int32_t)usens.size(), obs_usens.data(), Rm.data(), ucen.data());
      }
      if (rc == B200SFM_OK) rc = b200sfm_gp_problem_set_state(prob, centers.data(), points.data(), scales.data());
      if (rc == B200SFM_OK) rc = b200sfm_gp_problem_solve(prob, &o, &summary);
      if (rc == B200SFM_OK) rc = b200sfm_gp_problem_get_state(prob, centers.data(), points.data(), scales.data());
      if (rc == B200SFM_OK && !usens.empty()) {
        rc = b200sfm_gp_problem_get_rig_unknown(prob, ucen.data());
        if (rc == B200SFM_OK)
          for (auto& [key, idx] : usens) {   // ConvertResults: centre -> translation = -(R_cr u)  (.cc:578-582)
            Rigid3d cfr = b200host_adapt::CamFromRig(rigs[key.first], key.second);
            double Rs[9];
            QuatToR(cfr.rotation.coeffs().data(), Rs);
            for (int k = 0; k < 3; ++k)
              cfr.translation[k] = -(Rs[3 * k] * ucen[3 * idx] + Rs[3 * k + 1] * ucen[3 * idx + 1] + Rs[3 * k + 2] * ucen[3 * idx + 2]);
            b200host_adapt::SetCamFromRig(rigs[key.first], key.second, cfr);
          }
      }
      b200sfm_gp_problem_free(prob);
    } else {
      rc = b200sfm_gp_solve(ctx, &o, C, P, (int64_t)obs_cam.size(), ptb.data(), obs_cam.data(), obs_dir.data(),
                            calibrated.data(), nullptr, centers.data(), points.data(), scales.data(), &summary);
    }
    if (rc != B200SFM_OK) { std::fprintf(stderr, "b200sfm_gp_solve: %s\n", b200sfm_last_error(ctx)); return false; }
    for (auto& [id, f] : fsorted) {                                                           // ConvertResults: t = -R c (.cc:566-568)
      const int i = fidx[id];
      const double* R = &Rm[9 * (size_t)i];
      for (int k = 0; k < 3; ++k)
        f->RigFromWorld().translation[k] = -(R[3 * k] * centers[3 * i] + R[3 * k + 1] * centers[3 * i + 1] + R[3 * k + 2] * centers[3 * i + 2]);
    }
    p = 0;
    for (auto& [id, t] : tsorted) { for (int k = 0; k < 3; ++k) t->xyz[k] = points[3 * (size_t)p + k]; ++p; }
    return summary.usable != 0;
  }

 private:
  GlobalPositionerOptions options_;
  std::mt19937 random_generator_;
};

// ---------------------------------------------------------------------------
struct RotationEstimatorOptions {   // global_rotation_averaging.h:39-75
  int max_num_l1_iterations = 5;
  double l1_step_convergence_threshold = 0.001;
  int max_num_irls_iterations = 100;
  double irls_step_convergence_threshold = 0.001;
  double irls_loss_parameter_sigma = 5.0;
  enum WeightType { GEMAN_MCCLURE, HALF_NORM } weight_type = GEMAN_MCCLURE;
  bool skip_initialization = false;
  bool use_weight = false;
  bool use_gravity = false;
  double pcg_rel_tolerance = 1e-8;
};

class RotationEstimator {
 public:
  explicit RotationEstimator(const RotationEstimatorOptions& options) : options_(options) {}
  b200sfm_ra_stats summary{};

  // InitializeFromMaximumSpanningTree (global_rotation_averaging.cc:87-138 + math/tree.cc:78-153): Kruskal on
  // (max #inliers - #inliers), BFS from the first registered image, composition of the relative rotations along the
  // tree, then ConvertRotationsFromImageToRig (rotation_initializer.cc:7-120) for trivial frames and calibrated rigs:
  // rig_from_world = average over the frame's images of cam_from_rig^-1 * cam_from_world.  Images are enumerated in
  // sorted-id order (the reference: unordered_map order), so the root is the smallest registered image id.
  void InitializeFromMaximumSpanningTree(const ViewGraph& view_graph, std::unordered_map<rig_t, Rig>& rigs,
                                         std::unordered_map<frame_t, Frame>& frames,
                                         std::unordered_map<image_t, Image>& images) {
    auto registered = [&](const Image& im) {
      auto f = frames.find(im.frame_id);
      return f != frames.end() && f->second.is_registered;
    };
    std::map<image_t, int> idx;
    std::vector<image_t> ids;
    {
      std::map<image_t, const Image*> isorted;
      for (const auto& [id, im] : images) isorted[id] = &im;
      for (const auto& [id, im] : isorted)
        if (registered(*im)) { idx[id] = (int)ids.size(); ids.push_back(id); }
    }
    const int n = (int)ids.size();
    if (n == 0) return;
    struct Edge { double w; int a, b; const ImagePair* pr; };
    std::map<image_pair_t, const ImagePair*> psorted;
    double max_w = 0;
    for (const auto& [id, pr] : view_graph.image_pairs)
      if (pr.is_valid) { psorted[id] = &pr; max_w = std::max(max_w, (double)pr.inliers.size()); }   // tree.cc:93-100 (INLIER_NUM)
    std::vector<Edge> edges;
    for (const auto& [id, pr] : psorted) {
      auto a = idx.find(pr->image_id1), b = idx.find(pr->image_id2);
      if (a == idx.end() || b == idx.end()) continue;                                         // tree.cc:113-116
      edges.push_back({max_w - (double)pr->inliers.size(), a->second, b->second, pr});
    }
    std::stable_sort(edges.begin(), edges.end(), [](const Edge& x, const Edge& y) { return x.w < y.w; });
    std::vector<int> uf(n);
    for (int i = 0; i < n; ++i) uf[i] = i;
    auto find = [&](int x) { while (uf[x] != x) { uf[x] = uf[uf[x]]; x = uf[x]; } return x; };
    std::vector<std::vector<std::pair<int, const ImagePair*>>> adj(n);
    for (const Edge& e : edges) {
      const int ra = find(e.a), rb = find(e.b);
      if (ra == rb) continue;
      uf[ra] = rb;
      adj[e.a].push_back({e.b, e.pr});
      adj[e.b].push_back({e.a, e.pr});
    }
    // BFS from index 0; cam_from_world rotation of the root = identity (default-constructed, .cc:112)
    std::vector<std::array<double, 4>> q(n, {{0, 0, 0, 1}});
    std::vector<char> seen(n, 0);
    std::queue<int> bfs;
    bfs.push(0);
    seen[0] = 1;
    while (!bfs.empty()) {
      const int cur = bfs.front();
      bfs.pop();
      for (const auto& [nb, pr] : adj[cur]) {
        if (seen[nb]) continue;
        seen[nb] = 1;
        const double* r21 = pr->cam2_from_cam1.rotation.coeffs().data();
        if (pr->image_id1 == ids[nb]) {          // 1_R_w = 2_R_1^T * 2_R_w   (.cc:125-129)
          double inv[4];
          QuatConj(r21, inv);
          QuatMul(inv, q[cur].data(), q[nb].data());
        } else {                                 // 2_R_w = 2_R_1 * 1_R_w     (.cc:130-134)
          QuatMul(r21, q[cur].data(), q[nb].data());
        }
        bfs.push(nb);
      }
    }
    // ConvertRotationsFromImageToRig: per frame, average cam_from_rig^-1 * cam_from_world over its estimated images
    std::map<frame_t, std::vector<std::array<double, 4>>> per_frame;
    for (int i = 0; i < n; ++i) {
      if (!seen[i]) continue;                    // not reached by the tree: not estimated (rotation_initializer.cc:101-102)
      const Image& im = images.at(ids[i]);
      std::array<double, 4> r = q[i];
      if (!im.HasTrivialFrame()) {
        Rig& rig = rigs[frames[im.frame_id].RigId()];
        if (!b200host_adapt::IsRefSensor(rig, im.camera_id)) {
          if (!b200host_adapt::HasCamFromRig(rig, im.camera_id)) continue;                    // .cc:108-111
          const Rigid3d c = b200host_adapt::CamFromRig(rig, im.camera_id);
          double inv[4];
          QuatConj(c.rotation.coeffs().data(), inv);
          QuatMul(inv, q[i].data(), r.data());
        }
      }
      per_frame[im.frame_id].push_back(r);
    }
    for (auto& [fid, qs] : per_frame) {
      double avg[4];
      AverageQuaternions(qs, avg);
      double* out = frames[fid].RigFromWorld().rotation.coeffs().data();
      for (int k = 0; k < 4; ++k) out[k] = avg[k];
    }
  }

  // global_rotation_averaging.cc:40-85 (3-DoF frames and, with use_gravity, 1-DoF frames that carry a gravity
  // prior; trivial frames, known rigs and rigs with sensors whose cam_from_rig rotation is estimated alongside).
  bool EstimateRotations(const ViewGraph& view_graph, std::unordered_map<rig_t, Rig>& rigs,
                         std::unordered_map<frame_t, Frame>& frames, std::unordered_map<image_t, Image>& images) {
    if (options_.use_gravity) {   // .cc:47-59: gravity-aligned averaging needs every rig calibrated
      for (auto& [rig_id, rig] : rigs)
        if (!b200host_adapt::AllSensorsCalibrated(rig)) {
          std::fprintf(stderr, "Rig %u has an uncalibrated sensor, but the gravity aligned rotation is requested. "
                               "Please add the rig calibration.\n", (unsigned)rig_id);
          return false;
        }
    }
    b200sfm_ctx* ctx = DefaultContext();
    if (!ctx) return false;
    if (!options_.skip_initialization && !options_.use_gravity)                               // .cc:60-63
      InitializeFromMaximumSpanningTree(view_graph, rigs, frames, images);
    std::map<frame_t, Frame*> fsorted;
    for (auto& [id, f] : frames)
      if (f.is_registered) fsorted[id] = &f;
    std::map<frame_t, int> fidx;
    for (auto& [id, f] : fsorted) { const int i = (int)fidx.size(); fidx[id] = i; }
    const int n = (int)fsorted.size();
    if (n == 0) return false;
    std::vector<double> theta(3 * (size_t)n);
    for (auto& [id, f] : fsorted) QuatToAngleAxis(f->RigFromWorld().rotation.coeffs().data(), &theta[3 * (size_t)fidx[id]]);   // .cc:223-224
    // Cameras whose cam_from_rig rotation has to be estimated (.cc:162-194): non-reference sensors of the rigs of the
    // registered images without a cam_from_rig, or with one whose translation is still NaN (its rotation is then the
    // initial value, .cc:186-190; else zero, .cc:239-241).  They become nodes n, n + 1, ... in ascending camera id.
    std::map<camera_t, rig_t> cam_rig;
    for (auto& [id, im] : images) {
      const auto fit = frames.find(im.frame_id);
      if (fit == frames.end() || !fit->second.is_registered) continue;
      if (im.HasTrivialFrame()) continue;   // == IsRefSensor of its rig (scene/image.h:73-76): never estimated
      cam_rig[im.camera_id] = fit->second.RigId();
    }
    std::map<camera_t, int> ucam;
    for (auto& [cam, rig_id] : cam_rig) {
      Rig& rig = rigs[rig_id];
      if (b200host_adapt::IsRefSensor(rig, cam)) continue;
      const bool has = b200host_adapt::HasCamFromRig(rig, cam);
      bool nan_t = false;
      if (has) {
        const Rigid3d c = b200host_adapt::CamFromRig(rig, cam);
        for (int k = 0; k < 3; ++k) nan_t = nan_t || std::isnan(c.translation[k]);
      }
      if (!has || nan_t) {
        const int node = n + (int)ucam.size();
        ucam[cam] = node;
        double aa[3] = {0, 0, 0};
        if (has) {
          const Rigid3d c = b200host_adapt::CamFromRig(rig, cam);
          QuatToAngleAxis(c.rotation.coeffs().data(), aa);
        }
        theta.insert(theta.end(), aa, aa + 3);
      }
    }
    const int n_cams = (int)ucam.size();
    // use_gravity (.cc:207-217): a frame with a gravity prior keeps theta = (0, phi, 0), phi = RotUpToAngle(R_align^T R);
    // the first such frame (sorted-id order) is the fixed one
    std::vector<uint8_t> has_gravity(n, 0);
    std::vector<double> R_align(9 * (size_t)n, 0.0);
    int fixed_frame = 0;
    bool any_gravity = false;
    if (options_.use_gravity) {
      for (auto& [id, f] : fsorted) {
        if (!f->HasGravity()) continue;
        const int i = fidx[id];
        double* Ra = &R_align[9 * (size_t)i];
        b200host_adapt::RAlignRowMajor(*f, Ra);
        double R0[9], M[9], q[4], aa[3];
        QuatToR(f->RigFromWorld().rotation.coeffs().data(), R0);
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) M[3 * r + c] = Ra[r] * R0[c] + Ra[3 + r] * R0[3 + c] + Ra[6 + r] * R0[6 + c];   // R_align^T R
        RToQuat(M, q);
        QuatToAngleAxis(q, aa);
        theta[3 * (size_t)i] = 0.0; theta[3 * (size_t)i + 1] = aa[1]; theta[3 * (size_t)i + 2] = 0.0;
        has_gravity[i] = 1;
        if (!any_gravity) fixed_frame = i;
        any_gravity = true;
      }
    }
    std::map<image_pair_t, const ImagePair*> psorted;
    for (const auto& [id, pr] : view_graph.image_pairs)
      if (pr.is_valid) psorted[id] = &pr;
    std::vector<int32_t> ei, ej, eci, ecj;
    std::vector<double> Rrel, w;
    for (const auto& [id, pr] : psorted) {
      const auto i1 = images.find(pr->image_id1), i2 = images.find(pr->image_id2);
      if (i1 == images.end() || i2 == images.end()) continue;
      const auto f1 = fidx.find(i1->second.frame_id), f2 = fidx.find(i2->second.frame_id);
      if (f1 == fidx.end() || f2 == fidx.end()) continue;                                     // .cc:365-368
      double R[9];
      QuatToR(pr->cam2_from_cam1.rotation.coeffs().data(), R);
      // known rigs: the unknowns are the frame rotations, R_rel = R_c2r2^T R_21 R_c1r1 (.cc:274-309); an image
      // pair inside one frame is a self loop and is skipped (.cc:300-303)
      // (has_sensor_from_rig of the reference: a non-reference sensor with a KNOWN cam_from_rig; an unknown one
      //  contributes the identity here and its own -I / +I block, .cc:281-296,425-440)
      const auto u1 = ucam.find(i1->second.camera_id), u2 = ucam.find(i2->second.camera_id);
      const bool rig1 = !i1->second.HasTrivialFrame() && u1 == ucam.end();
      const bool rig2 = !i2->second.HasTrivialFrame() && u2 == ucam.end();
      if (rig1 && rig2 && f1->second == f2->second) continue;
      if (rig1) {
        const Rigid3d c = b200host_adapt::CamFromRig(rigs[frames[i1->second.frame_id].RigId()], i1->second.camera_id);
        double Rc[9], T[9];
        QuatToR(c.rotation.coeffs().data(), Rc);
        for (int r = 0; r < 3; ++r)
          for (int k = 0; k < 3; ++k) T[3 * r + k] = R[3 * r] * Rc[k] + R[3 * r + 1] * Rc[3 + k] + R[3 * r + 2] * Rc[6 + k];
        std::copy(T, T + 9, R);
      }
      if (rig2) {
        const Rigid3d c = b200host_adapt::CamFromRig(rigs[frames[i2->second.frame_id].RigId()], i2->second.camera_id);
        double Rc[9], T[9];
        QuatToR(c.rotation.coeffs().data(), Rc);
        for (int r = 0; r < 3; ++r)   // Rc^T R
          for (int k = 0; k < 3; ++k) T[3 * r + k] = Rc[r] * R[k] + Rc[3 + r] * R[3 + k] + Rc[6 + r] * R[6 + k];
        std::copy(T, T + 9, R);
      }
      if (any_gravity) {   // align the relative rotation with the gravity frames (.cc:311-326)
        double T[9];
        if (has_gravity[f1->second]) {
          const double* Ra = &R_align[9 * (size_t)f1->second];
          for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 3; ++k) T[3 * r + k] = R[3 * r] * Ra[k] + R[3 * r + 1] * Ra[3 + k] + R[3 * r + 2] * Ra[6 + k];
          std::copy(T, T + 9, R);
        }
        if (has_gravity[f2->second]) {
          const double* Ra = &R_align[9 * (size_t)f2->second];
          for (int r = 0; r < 3; ++r)   // R_align^T R
            for (int k = 0; k < 3; ++k) T[3 * r + k] = Ra[r] * R[k] + Ra[3 + r] * R[3 + k] + Ra[6 + r] * R[6 + k];
          std::copy(T, T + 9, R);
        }
      }
      ei.push_back(f1->second);
      ej.push_back(f2->second);
      eci.push_back(u1 == ucam.end() || i1->second.HasTrivialFrame() ? -1 : u1->second);
      ecj.push_back(u2 == ucam.end() || i2->second.HasTrivialFrame() ? -1 : u2->second);
      Rrel.insert(Rrel.end(), R, R + 9);
      w.push_back(pr->weight);
    }
    b200sfm_ra_opts o;
    b200sfm_ra_default_opts(&o);
    o.max_num_l1_iterations = options_.max_num_l1_iterations; o.max_num_irls_iterations = options_.max_num_irls_iterations;
    o.l1_step_convergence_threshold = options_.l1_step_convergence_threshold;
    o.irls_step_convergence_threshold = options_.irls_step_convergence_threshold;
    o.irls_loss_parameter_sigma = options_.irls_loss_parameter_sigma;
    o.weight_type = options_.weight_type == RotationEstimatorOptions::HALF_NORM ? 1 : 0;
    o.use_weight = options_.use_weight; o.pcg_rel_tolerance = options_.pcg_rel_tolerance;
    int rc;
    if (n_cams > 0) {   // frames that hold an image of each unknown camera (the quaternion average of .cc:675-693 runs over them)
      std::map<camera_t, std::set<int>> cam_frames_of;
      for (auto& [id, im] : images) {
        const auto u = ucam.find(im.camera_id);
        const auto f = fidx.find(im.frame_id);
        if (u != ucam.end() && f != fidx.end() && !im.HasTrivialFrame()) cam_frames_of[im.camera_id].insert(f->second);
      }
      std::vector<int32_t> cfb(1, 0), cf;
      for (auto& [cam, node] : ucam) {
        for (int f : cam_frames_of[cam]) cf.push_back(f);
        cfb.push_back((int32_t)cf.size());
      }
      rc = b200sfm_ra_solve_rig(ctx, &o, n, n_cams, (int64_t)ei.size(), ei.data(), ej.data(), eci.data(), ecj.data(), Rrel.data(),
                                w.data(), cfb.data(), cf.data(), 0, theta.data(), &summary);
    } else if (any_gravity) {
      rc = b200sfm_ra_solve_gravity(ctx, &o, n, (int64_t)ei.size(), ei.data(), ej.data(), Rrel.data(), w.data(), has_gravity.data(),
                                    fixed_frame, theta.data(), &summary);
    } else {
      rc = b200sfm_ra_solve(ctx, &o, n, (int64_t)ei.size(), ei.data(), ej.data(), Rrel.data(), w.data(), 0, theta.data(), &summary);
    }
    if (rc != B200SFM_OK) { std::fprintf(stderr, "b200sfm_ra_solve: %s\n", b200sfm_last_error(ctx)); return false; }
    if (!summary.usable) return false;                                                        // NaN (.cc:508-512,590-593)
    for (auto& [id, f] : fsorted) {                                                           // ConvertResults (.cc:787-798)
      const int i = fidx[id];
      double* qout = f->RigFromWorld().rotation.coeffs().data();
      if (has_gravity[i]) {   // R = R_align * AngleToRotUp(phi)
        double qa[4], Ry[9], M[9];
        AngleAxisToQuat(&theta[3 * (size_t)i], qa);
        QuatToR(qa, Ry);
        const double* Ra = &R_align[9 * (size_t)i];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) M[3 * r + c] = Ra[3 * r] * Ry[c] + Ra[3 * r + 1] * Ry[3 + c] + Ra[3 * r + 2] * Ry[6 + c];
        RToQuat(M, qout);
      } else {
        AngleAxisToQuat(&theta[3 * (size_t)i], qout);
      }
      for (int k = 0; k < 3; ++k) f->RigFromWorld().translation[k] = 0.0;                     // Vector3d::Zero() (.cc:795)
    }
    for (auto& [cam, node] : ucam) {   // the estimated cam_from_rig rotations, translation not known yet (.cc:800-813)
      Rigid3d c;
      AngleAxisToQuat(&theta[3 * (size_t)node], c.rotation.coeffs().data());
      for (int k = 0; k < 3; ++k) c.translation[k] = std::numeric_limits<double>::quiet_NaN();
      b200host_adapt::SetCamFromRig(rigs[cam_rig[cam]], cam, c);
    }
    return true;
  }

  // rotation matrix (row-major) -> unit quaternion xyzw, w >= 0 (Shepperd's branches, as Eigen's conversion)
  static void RToQuat(const double* R, double* q) {
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
      const double s = std::sqrt(tr + 1.0) * 2;
      q[3] = 0.25 * s; q[0] = (R[7] - R[5]) / s; q[1] = (R[2] - R[6]) / s; q[2] = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
      const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
      q[3] = (R[7] - R[5]) / s; q[0] = 0.25 * s; q[1] = (R[1] + R[3]) / s; q[2] = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
      const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
      q[3] = (R[2] - R[6]) / s; q[0] = (R[1] + R[3]) / s; q[1] = 0.25 * s; q[2] = (R[5] + R[7]) / s;
    } else {
      const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
      q[3] = (R[3] - R[1]) / s; q[0] = (R[2] + R[6]) / s; q[1] = (R[5] + R[7]) / s; q[2] = 0.25 * s;
    }
    if (q[3] < 0) for (int k = 0; k < 4; ++k) q[k] = -q[k];
  }

  static void QuatToAngleAxis(const double* q, double* v) {
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (n > 0) {
      const double ang = 2 * std::atan2(n, std::fabs(q[3]));
      const double f = (q[3] < 0 ? -ang : ang) / n;
      v[0] = q[0] * f; v[1] = q[1] * f; v[2] = q[2] * f;
    } else {
      v[0] = v[1] = v[2] = 0;
    }
  }
  static void AngleAxisToQuat(const double* v, double* q) {
    const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (n > 0) {
      const double s = std::sin(n / 2) / n;
      q[0] = v[0] * s; q[1] = v[1] * s; q[2] = v[2] * s; q[3] = std::cos(n / 2);
    } else {
      q[0] = q[1] = q[2] = 0; q[3] = 1;
    }
  }

 private:
  const RotationEstimatorOptions& options_;   // the reference stores a const& too (.h:140)
};

}  // namespace b200sfm_shim
