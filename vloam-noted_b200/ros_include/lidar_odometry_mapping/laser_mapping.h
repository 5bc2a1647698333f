// Forwarding header: takes the place of src/lidar_odometry_mapping/include/lidar_odometry_mapping/laser_mapping.h of the reference.
// The four stage classes (same names, constructors and member signatures) live in vloam_adapter.hpp on top of the C ABI.
#pragma once
#include "vloam_adapter.hpp"
