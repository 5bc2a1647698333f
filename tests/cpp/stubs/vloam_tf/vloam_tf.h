// stand-in for src/vloam_tf/include/vloam_tf/vloam_tf.h: the transform blackboard shared by VO / LO / MO.  Only the
// members the lidar path reads or writes are declared (vloam_tf.h:25-50), with the reference's names and types.
#pragma once
#include <tf2/LinearMath/Transform.h>
namespace vloam {
class VloamTF {
public:
  void init() { base_T_cam0.setIdentity(); velo_last_VOT_velo_curr.setIdentity(); }  // (the real one also opens the tf listener)
  tf2::Transform base_T_cam0;                                    // static, written by processStaticTransform()
  tf2::Transform velo_last_VOT_velo_curr;                        // VO prior read by LaserOdometry when detach_VO_LO == false
  tf2::Transform world_LOT_base_last, base_prev_LOT_base_curr, cam0_curr_LOT_cam0_prev;  // written by LaserOdometry::publish
  tf2::Transform world_MOT_base_last;                            // written by LaserMapping::publish
};
}  // namespace vloam
