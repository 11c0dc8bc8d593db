// TEST STUB (tests only, never shipped): just enough of the public API of Eigen / COLMAP / glomap -- with the REAL
// spellings and types (Eigen::Quaterniond::coeffs().data(), Eigen::Vector3d, enum class CameraModelId, sensor_t,
// std::optional<Rigid3d> MaybeSensorFromRig, Frame::is_registered, ...) -- for
//   g++ -fsyntax-only -DB200SFM_WITH_GLOMAP -I tests/shim_mock/glomap_stub glomap_b200/host/estimators_shim.h
// to type-check the branch of the shim that is compiled inside a glomap build.  Signatures
// follow glomap/scene/{types.h,image.h,frame.h,track.h,camera.h,image_pair.h,view_graph.h} at 99806d0 and the COLMAP
// headers those include (colmap/geometry/rigid3.h, colmap/scene/rig.h, colmap/sensor/models.h).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <optional>
#include <set>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace Eigen {
template <int N>
struct VecN {
  double v[N] = {};
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
  double& operator()(int i) { return v[i]; }
  const double& operator()(int i) const { return v[i]; }
  double* data() { return v; }
  const double* data() const { return v; }
  bool hasNaN() const {
    for (double x : v)
      if (std::isnan(x)) return true;
    return false;
  }
  void setConstant(double c) {
    for (double& x : v) x = c;
  }
  static VecN Zero() { return VecN(); }
};
using Vector2d = VecN<2>;
using Vector3d = VecN<3>;
struct Vector4dMap {
  double* p;
  double* data() const { return p; }
};
struct ConstVector4dMap {
  const double* p;
  const double* data() const { return p; }
};
struct Quaterniond {
  double c[4] = {0, 0, 0, 1};   // x y z w
  Vector4dMap coeffs() { return Vector4dMap{c}; }
  ConstVector4dMap coeffs() const { return ConstVector4dMap{c}; }
};
struct Matrix3d {
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double operator()(int r, int c) const { return m[3 * r + c]; }
};
struct VectorXi {
  std::vector<int> v;
  long size() const { return (long)v.size(); }
};
}  // namespace Eigen

namespace colmap {
enum class CameraModelId { kInvalid = -1, kSimplePinhole = 0, kPinhole = 1, kSimpleRadial = 2, kRadial = 3 };
enum class SensorType { INVALID = -1, CAMERA = 0, IMU = 1 };
struct sensor_t {
  SensorType type;
  uint32_t id;
  constexpr sensor_t(SensorType t = SensorType::INVALID, uint32_t i = 0) : type(t), id(i) {}
  bool operator<(const sensor_t& o) const { return std::make_pair((int)type, id) < std::make_pair((int)o.type, o.id); }
};
struct data_t {
  sensor_t sensor_id;
  uint32_t id;
};
struct Rigid3d {
  Eigen::Quaterniond rotation;
  Eigen::Vector3d translation;
};
class Rig {
 public:
  sensor_t RefSensorId() const { return ref_; }
  bool IsRefSensor(sensor_t s) const { return s.id == ref_.id && s.type == ref_.type; }
  const std::map<sensor_t, std::optional<Rigid3d>>& NonRefSensors() const { return non_ref_; }
  std::map<sensor_t, std::optional<Rigid3d>>& NonRefSensors() { return non_ref_; }
  Rigid3d& SensorFromRig(sensor_t s) { return *non_ref_.at(s); }
  const Rigid3d& SensorFromRig(sensor_t s) const { return *non_ref_.at(s); }
  const std::optional<Rigid3d>& MaybeSensorFromRig(sensor_t s) const { return non_ref_.at(s); }
  void SetSensorFromRig(sensor_t s, const Rigid3d& r) { non_ref_[s] = r; }

 private:
  sensor_t ref_;
  std::map<sensor_t, std::optional<Rigid3d>> non_ref_;
};
}  // namespace colmap

namespace glomap {
using colmap::Rig;
using colmap::Rigid3d;
using colmap::sensor_t;
using colmap::SensorType;
using image_t = uint32_t;
using camera_t = uint32_t;
using frame_t = uint32_t;
using rig_t = uint32_t;
using track_t = uint64_t;
using image_pair_t = uint64_t;
using feature_t = uint32_t;

struct Camera {   // glomap/scene/camera.h: colmap::Camera + has_prior_focal_length
  camera_t camera_id = 0;
  colmap::CameraModelId model_id = colmap::CameraModelId::kInvalid;
  std::vector<double> params;
  bool has_prior_focal_length = false;
};
struct GravityInfo {   // glomap/scene/frame.h:11-27
  bool has_gravity = false;
  const Eigen::Matrix3d& GetRAlign() const { return R_align_; }

 private:
  Eigen::Matrix3d R_align_;
};
struct Frame {   // glomap/scene/frame.h: colmap::Frame + is_registered, gravity_info
  GravityInfo gravity_info;
  bool is_registered = false;
  bool HasGravity() const { return gravity_info.has_gravity; }
  rig_t RigId() const { return rig_id_; }
  Rig* RigPtr() const { return rig_ptr_; }
  bool HasPose() const { return rig_from_world_.has_value(); }
  Rigid3d& RigFromWorld() { return *rig_from_world_; }
  const Rigid3d& RigFromWorld() const { return *rig_from_world_; }
  const std::optional<Rigid3d>& MaybeRigFromWorld() const { return rig_from_world_; }
  void SetRigFromWorld(const Rigid3d& r) { rig_from_world_ = r; }
  const std::set<colmap::data_t>& DataIds() const { return data_ids_; }

 private:
  rig_t rig_id_ = 0;
  Rig* rig_ptr_ = nullptr;
  std::optional<Rigid3d> rig_from_world_;
  std::set<colmap::data_t> data_ids_;
};
struct Image {   // glomap/scene/image.h:10-45
  image_t image_id = 0;
  camera_t camera_id = 0;
  frame_t frame_id = 0;
  std::string file_name;
  Frame* frame_ptr = nullptr;
  std::vector<Eigen::Vector2d> features;
  std::vector<Eigen::Vector3d> features_undist;
  bool IsRegistered() const { return frame_ptr != nullptr && frame_ptr->is_registered; }
  bool HasTrivialFrame() const { return trivial_; }
  bool HasGravity() const { return frame_ptr->HasGravity(); }

 private:
  bool trivial_ = true;
};
struct Track {   // glomap/scene/track.h:12-24
  track_t track_id = 0;
  Eigen::Vector3d xyz;
  std::vector<std::pair<image_t, feature_t>> observations;
  bool is_initialized = false;
};
struct ImagePair {   // glomap/scene/image_pair.h:13-40
  image_t image_id1 = 0, image_id2 = 0;
  bool is_valid = true;
  double weight = -1;
  Eigen::VectorXi inliers;
  Rigid3d cam2_from_cam1;
};
struct ViewGraph {   // glomap/scene/view_graph.h:12
  std::unordered_map<image_pair_t, ImagePair> image_pairs;
};
}  // namespace glomap
namespace colmap {
inline bool operator<(const data_t& a, const data_t& b) { return a.id < b.id; }
}
