// stand-in for <tf2/LinearMath/Transform.h>: rigid transforms as the VloamTF fields use them
// (setOrigin / setRotation / getOrigin / getRotation / inverse / operator*: laser_odometry.cpp:612-620, laser_mapping.cpp:834-861)
#pragma once
#include <math.h>
namespace tf2 {
class Vector3 {
public:
  Vector3() {}
  Vector3(double x, double y, double z) { v_[0] = x; v_[1] = y; v_[2] = z; }
  double x() const { return v_[0]; } double y() const { return v_[1]; } double z() const { return v_[2]; }
private:
  double v_[3] = {0, 0, 0};
};
class Quaternion {
public:
  Quaternion() {}
  Quaternion(double x, double y, double z, double w) { q_[0] = x; q_[1] = y; q_[2] = z; q_[3] = w; }  // tf2's (x, y, z, w) order
  double x() const { return q_[0]; } double y() const { return q_[1]; } double z() const { return q_[2]; } double w() const { return q_[3]; }
  Quaternion inverse() const { return Quaternion(-q_[0], -q_[1], -q_[2], q_[3]); }
  Quaternion operator*(const Quaternion& b) const {
    return Quaternion(q_[3] * b.q_[0] + q_[0] * b.q_[3] + q_[1] * b.q_[2] - q_[2] * b.q_[1], q_[3] * b.q_[1] + q_[1] * b.q_[3] + q_[2] * b.q_[0] - q_[0] * b.q_[2],
                      q_[3] * b.q_[2] + q_[2] * b.q_[3] + q_[0] * b.q_[1] - q_[1] * b.q_[0], q_[3] * b.q_[3] - q_[0] * b.q_[0] - q_[1] * b.q_[1] - q_[2] * b.q_[2]);
  }
  Vector3 rotate(const Vector3& v) const {
    const Quaternion p(v.x(), v.y(), v.z(), 0), r = (*this) * p * inverse();
    return Vector3(r.x(), r.y(), r.z());
  }
private:
  double q_[4] = {0, 0, 0, 1};
};
class Transform {
public:
  Transform() {}
  Transform(const Quaternion& q, const Vector3& o) : q_(q), o_(o) {}
  void setOrigin(const Vector3& o) { o_ = o; }
  void setRotation(const Quaternion& q) { q_ = q; }
  void setIdentity() { q_ = Quaternion(); o_ = Vector3(); }
  const Vector3& getOrigin() const { return o_; }
  Quaternion getRotation() const { return q_; }
  Transform inverse() const { const Quaternion qi = q_.inverse(); const Vector3 t = qi.rotate(o_); return Transform(qi, Vector3(-t.x(), -t.y(), -t.z())); }
  Transform operator*(const Transform& b) const {
    const Vector3 r = q_.rotate(b.o_);
    return Transform(q_ * b.q_, Vector3(r.x() + o_.x(), r.y() + o_.y(), r.z() + o_.z()));
  }
private:
  Quaternion q_;
  Vector3 o_;
};
}  // namespace tf2
