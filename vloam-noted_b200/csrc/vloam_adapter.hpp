// vloam_adapter.hpp -- the reference's C++ stage classes re-created on top of the C ABI.
//
// Drop-in for src/lidar_odometry_mapping/include/lidar_odometry_mapping/
//   {scan_registration.h, laser_odometry.h, laser_mapping.h, lidar_odometry_mapping.h}:
// same class names, member functions, argument order and meaning, so the only caller
// (vloam_main_node.cpp:143-144, 186-190 through LidarOdometryMapping) compiles unchanged.
//
// With -DVLOAM_ADAPTER_WITH_PCL (a ROS/PCL/Eigen box) the signatures use pcl::PointCloud
// and Eigen types exactly like the reference.  Without it (this repo's test box has
// neither) the same classes are built on the small stand-in types below so the adapter
// logic can be compiled and tested; the marshalling (PCL's 32-byte PointXYZI <-> the
// ABI's 16-byte x,y,z,intensity) is the only difference.
//
// publish() of the three classes (SR.cpp:516-564, LO.cpp:587-657, LM.cpp:816-920) is ROS
// message plumbing and stays with the caller; poses are available from output().
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "vloam_b200.h"

#ifdef VLOAM_ADAPTER_WITH_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <Eigen/Dense>
namespace vloam {
typedef pcl::PointXYZI PointType;  // common.h:42
typedef pcl::PointCloud<pcl::PointXYZ> CloudXYZ;
typedef pcl::PointCloud<PointType> CloudXYZI;
typedef CloudXYZI::Ptr CloudPtr;
typedef Eigen::Quaterniond Quat;
typedef Eigen::Vector3d Vec3;
inline CloudPtr make_cloud() { return CloudPtr(new CloudXYZI()); }
inline void quat_set(Quat& q, const double* v) { q = Quat(v[3], v[0], v[1], v[2]); }
inline void vec_set(Vec3& t, const double* v) { t = Vec3(v[0], v[1], v[2]); }
inline void quat_get(const Quat& q, double* v) { v[0] = q.x(); v[1] = q.y(); v[2] = q.z(); v[3] = q.w(); }
inline void vec_get(const Vec3& t, double* v) { v[0] = t.x(); v[1] = t.y(); v[2] = t.z(); }
}  // namespace vloam
#else
namespace vloam {
struct PointXYZ { float x, y, z; };
struct PointType { float x, y, z, intensity; };
template <typename P> struct Cloud { std::vector<P> points; size_t size() const { return points.size(); } void clear() { points.clear(); } };
typedef Cloud<PointXYZ> CloudXYZ;
typedef Cloud<PointType> CloudXYZI;
typedef std::shared_ptr<CloudXYZI> CloudPtr;
struct Quat { double x = 0, y = 0, z = 0, w = 1; };
struct Vec3 { double x = 0, y = 0, z = 0; };
inline CloudPtr make_cloud() { return std::make_shared<CloudXYZI>(); }
inline void quat_set(Quat& q, const double* v) { q.x = v[0]; q.y = v[1]; q.z = v[2]; q.w = v[3]; }
inline void vec_set(Vec3& t, const double* v) { t.x = v[0]; t.y = v[1]; t.z = v[2]; }
inline void quat_get(const Quat& q, double* v) { v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
inline void vec_get(const Vec3& t, double* v) { v[0] = t.x; v[1] = t.y; v[2] = t.z; }
}  // namespace vloam
#endif

namespace vloam {

// The reference aborts (ROS_BREAK) on a missing parameter or a bad scan_line; the adapter throws.
struct AdapterError : std::runtime_error { using std::runtime_error::runtime_error; };

// Eigen's Quaterniond * Vector3d (uv = 2 u x v; v + w uv + u x uv), as the device code evaluates it.  Used only by
// the per-point helpers the reference exposes publicly (TransformToStart, pointAssociateToMap, ...): callers
// outside the pipeline may use them, the pipeline itself transforms whole clouds on the device.
inline void quat_rotate(const double q[4], const double v[3], double o[3]) {
  const double ux = q[0], uy = q[1], uz = q[2], w = q[3];
  double a = uy * v[2] - uz * v[1], b = uz * v[0] - ux * v[2], c = ux * v[1] - uy * v[0];
  a += a; b += b; c += c;
  o[0] = (v[0] + w * a) + (uy * c - uz * b);
  o[1] = (v[1] + w * b) + (uz * a - ux * c);
  o[2] = (v[2] + w * c) + (ux * b - uy * a);
}
inline void quat_inverse(const double q[4], double o[4]) {  // Eigen: conjugate / squaredNorm
  const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  o[0] = -q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = q[3] / n2;
}

class Engine {  // one vloam_b200_ctx shared by the three stage objects (they share it in LOM.h:78-80 too)
public:
  explicit Engine(const vloam_b200_params& p, int device = 0) {
    const int r = vloam_b200_create(&p, device, &ctx_);
    if (r != VLOAM_OK) throw AdapterError("vloam_b200_create failed (bad scan_line / resolution or no CUDA device): " + std::to_string(r));
  }
  ~Engine() { vloam_b200_destroy(ctx_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  vloam_b200_ctx* ctx() const { return ctx_; }
  void check(int r) const { if (r < 0) throw AdapterError(vloam_b200_last_error(ctx_)); }
  void fetch(int which, CloudPtr& out) const {
    const int n = vloam_b200_get_cloud(ctx_, which, nullptr, 0);
    check(n);
    std::vector<float> buf((size_t)n * 4);
    if (n) check(vloam_b200_get_cloud(ctx_, which, buf.data(), n));
    if (!out) out = make_cloud();
    out->points.resize(n);
    for (int i = 0; i < n; ++i) { PointType p; p.x = buf[i * 4]; p.y = buf[i * 4 + 1]; p.z = buf[i * 4 + 2]; p.intensity = buf[i * 4 + 3]; out->points[i] = p; }
  }
private:
  vloam_b200_ctx* ctx_ = nullptr;
};

class ScanRegistration {  // scan_registration.h:64-81
public:
  explicit ScanRegistration(std::shared_ptr<Engine> e) : e_(e) {}
  void init() {}    // parameters are bound when the Engine is created (SR.cpp:42-92)
  void reset() {}   // device buffers are overwritten by the next input() (SR.cpp:95-104)
  void input(const CloudXYZ& laserCloudIn_) {  // SR.cpp:144-513
    xyz_.resize(laserCloudIn_.points.size() * 3);
    for (size_t i = 0; i < laserCloudIn_.points.size(); ++i) { xyz_[i * 3] = laserCloudIn_.points[i].x; xyz_[i * 3 + 1] = laserCloudIn_.points[i].y; xyz_[i * 3 + 2] = laserCloudIn_.points[i].z; }
    e_->check(vloam_b200_scan_registration(e_->ctx(), xyz_.data(), (int)laserCloudIn_.points.size(), 3));
  }
  // SR.cpp:107-141 (public template of the reference): order-preserving removal of points closer than thres
  template <typename CloudT>
  static void removeClosedPointCloud(const CloudT& cloud_in, CloudT& cloud_out, float thres) {
    if (&cloud_in != &cloud_out) cloud_out.points.resize(cloud_in.points.size());
    size_t j = 0;
    for (size_t i = 0; i < cloud_in.points.size(); ++i) {
      const auto& p = cloud_in.points[i];
      if (p.x * p.x + p.y * p.y + p.z * p.z < thres * thres) continue;
      cloud_out.points[j++] = p;
    }
    cloud_out.points.resize(j);
  }
  void publish() {}
  void output(CloudPtr& laserCloud_, CloudPtr& cornerPointsSharp_, CloudPtr& cornerPointsLessSharp_, CloudPtr& surfPointsFlat_,
              CloudPtr& surfPointsLessFlat_) {  // SR.cpp:566-577
    e_->fetch(VLOAM_CLOUD_FULL, laserCloud_); e_->fetch(VLOAM_CLOUD_SHARP, cornerPointsSharp_);
    e_->fetch(VLOAM_CLOUD_LESS_SHARP, cornerPointsLessSharp_); e_->fetch(VLOAM_CLOUD_FLAT, surfPointsFlat_);
    e_->fetch(VLOAM_CLOUD_LESS_FLAT, surfPointsLessFlat_);
  }
private:
  std::shared_ptr<Engine> e_;
  std::vector<float> xyz_;
};

class LaserOdometry {  // laser_odometry.h:63-87
public:
  explicit LaserOdometry(std::shared_ptr<Engine> e) : e_(e) {}
  void init() {}
  // The five clouds already live on the device; input() is kept for source compatibility (LO.cpp:137-148).
  void input(const CloudPtr&, const CloudPtr&, const CloudPtr&, const CloudPtr&, const CloudPtr&) {}
  // detach_VO_LO == false: hand the VO prior velo_last_VOT_velo_curr over before solveLO (LO.cpp:237-250)
  void setPrior(const Quat& q, const Vec3& t) { quat_get(q, pq_); vec_get(t, pt_); use_prior_ = true; }
  void clearPrior() { use_prior_ = false; }
  void solveLO() {  // LO.cpp:199-584
    e_->check(vloam_b200_laser_odometry(e_->ctx(), pq_, pt_, use_prior_ ? 1 : 0, qw_, tw_, ql_, tl_, &skip_));
  }
  void publish() {}
  void output(Quat& q_w_curr_, Vec3& t_w_curr_, CloudPtr& laserCloudCornerLast_, CloudPtr& laserCloudSurfLast_, CloudPtr& laserCloudFullRes_,
              bool& skip_frame) {  // LO.cpp:660-679
    quat_set(q_w_curr_, qw_); vec_set(t_w_curr_, tw_);
    skip_frame = skip_ != 0;
    if (!skip_frame) { e_->fetch(VLOAM_CLOUD_CORNER_LAST, laserCloudCornerLast_); e_->fetch(VLOAM_CLOUD_SURF_LAST, laserCloudSurfLast_); e_->fetch(VLOAM_CLOUD_FULL, laserCloudFullRes_); }
  }
  void lastMotion(Quat& q_last_curr, Vec3& t_last_curr) const { quat_set(q_last_curr, ql_); vec_set(t_last_curr, tl_); }
  // LO.cpp:152-173 with DISTORTION == false (s = 1): the point of this sweep in the frame of the sweep's start
  void TransformToStart(PointType const* const pi, PointType* const po) const {
    const double v[3] = {pi->x, pi->y, pi->z};
    double r[3]; quat_rotate(ql_, v, r);
    po->x = (float)(r[0] + tl_[0]); po->y = (float)(r[1] + tl_[1]); po->z = (float)(r[2] + tl_[2]); po->intensity = pi->intensity;
  }
  // LO.cpp:176-193: into the frame of the sweep's end (dead in the reference: only called under `if (0)`, LO.cpp:537)
  void TransformToEnd(PointType const* const pi, PointType* const po) const {
    PointType un; TransformToStart(pi, &un);
    const double v[3] = {un.x - tl_[0], un.y - tl_[1], un.z - tl_[2]};
    double qi[4], r[3]; quat_inverse(ql_, qi); quat_rotate(qi, v, r);
    po->x = (float)r[0]; po->y = (float)r[1]; po->z = (float)r[2]; po->intensity = (float)(int)pi->intensity;
  }
private:
  std::shared_ptr<Engine> e_;
  double pq_[4] = {0, 0, 0, 1}, pt_[3] = {0, 0, 0}, qw_[4] = {0, 0, 0, 1}, tw_[3] = {0, 0, 0}, ql_[4] = {0, 0, 0, 1}, tl_[3] = {0, 0, 0};
  bool use_prior_ = false;
  int skip_ = 0;
};

class LaserMapping {  // laser_mapping.h:72-100
public:
  explicit LaserMapping(std::shared_ptr<Engine> e) : e_(e) {}
  void init() {}
  void reset() { e_->check(vloam_b200_begin_frame(e_->ctx())); }  // LM.cpp:132-136
  // The odometry outputs are read from the shared context (LM.cpp:178-209 happens inside laser_mapping).
  void input(const CloudPtr&, const CloudPtr&, const CloudPtr&, const Quat&, const Vec3&, const bool&) {}
  void solveMapping() { e_->check(vloam_b200_laser_mapping(e_->ctx(), q_, t_)); }  // LM.cpp:212-814 (skip frames: LM.cpp:197-201)
  void publish() {}
  void output(Quat& q_w_curr, Vec3& t_w_curr) const { quat_set(q_w_curr, q_); vec_set(t_w_curr, t_); }
  void transformUpdate() {}          // LM.cpp:147-151: already applied on the device at the end of solveMapping
  void transformAssociateToMap() {}  // declared in LM.h:87, its definition is commented out in the reference (LM.cpp:138-143)
  // LM.cpp:154-164 / 166-175 with the mapped pose of the last solveMapping
  void pointAssociateToMap(PointType const* const pi, PointType* const po) const {
    const double v[3] = {pi->x, pi->y, pi->z};
    double r[3]; quat_rotate(q_, v, r);
    po->x = (float)(r[0] + t_[0]); po->y = (float)(r[1] + t_[1]); po->z = (float)(r[2] + t_[2]); po->intensity = pi->intensity;
  }
  void pointAssociateTobeMapped(PointType const* const pi, PointType* const po) const {
    const double v[3] = {pi->x - t_[0], pi->y - t_[1], pi->z - t_[2]};
    double qi[4], r[3]; quat_inverse(q_, qi); quat_rotate(qi, v, r);
    po->x = (float)r[0]; po->y = (float)r[1]; po->z = (float)r[2]; po->intensity = pi->intensity;
  }
  // LM.cpp:901-905: the full-resolution cloud of this sweep in the map frame (on the device)
  void registeredFullCloud(CloudPtr& out) const {
    const int n = vloam_b200_register_full_cloud(e_->ctx(), nullptr, 0);
    e_->check(n);
    std::vector<float> buf((size_t)n * 4);
    if (n) e_->check(vloam_b200_register_full_cloud(e_->ctx(), buf.data(), n));
    if (!out) out = make_cloud();
    out->points.resize(n);
    for (int i = 0; i < n; ++i) { PointType p; p.x = buf[i * 4]; p.y = buf[i * 4 + 1]; p.z = buf[i * 4 + 2]; p.intensity = buf[i * 4 + 3]; out->points[i] = p; }
  }
private:
  std::shared_ptr<Engine> e_;
  double q_[4] = {0, 0, 0, 1}, t_[3] = {0, 0, 0};
};

class LidarOdometryMapping {  // lidar_odometry_mapping.h:45-86
public:
  explicit LidarOdometryMapping(const vloam_b200_params& p, int device = 0)
      : e_(std::make_shared<Engine>(p, device)), scan_registration(e_), laser_odometry(e_), laser_mapping(e_) {}
  void init() {}
  void reset() { scan_registration.reset(); laser_mapping.reset(); }                  // LOM.cpp:65-71
  void scanRegistrationIO(const CloudXYZ& laserCloudIn) { scan_registration.input(laserCloudIn); }  // LOM.cpp:77-100
  void laserOdometryIO() { laser_odometry.solveLO(); }                                // LOM.cpp:110-141
  void laserMappingIO() { laser_mapping.solveMapping(); }                             // LOM.cpp:144-176 (skip handled inside)
private:
  std::shared_ptr<Engine> e_;
public:
  ScanRegistration scan_registration;
  LaserOdometry laser_odometry;
  LaserMapping laser_mapping;
};

}  // namespace vloam
