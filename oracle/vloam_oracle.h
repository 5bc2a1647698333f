// vloam_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's lidar registration hot path
// (liuzm-slam/VLOAM-NOTED, src/lidar_odometry_mapping): scanRegistration ->
// laserOdometry -> laserMapping, plus the third-party semantics the reference
// relies on (PCL VoxelGrid, FLANN exact kNN, Ceres 2.0 trust-region LM with
// Huber loss and the Eigen quaternion manifold, Eigen quaternion algebra,
// 3x3 symmetric eigen-decomposition, 5x3 column-pivoted Householder QR).
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
// for this path and cannot be compiled here (ROS1, PCL, FLANN, Ceres, Eigen are
// absent), so this oracle is pinned only by its own
// self-checks (tests/test_oracle_*.py) and by the citations below.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may link or call this code.  The product (libvloam_b200.so)
// never does.
//
// File aliases in citations (relative to /root/reference/src/lidar_odometry_mapping):
//   SR.cpp = src/scan_registration.cpp      SR.h = include/.../scan_registration.h
//   LO.cpp = src/laser_odometry.cpp         LO.h = include/.../laser_odometry.h
//   LM.cpp = src/laser_mapping.cpp          LM.h = include/.../laser_mapping.h
//   LF.hpp = include/.../lidarFactor.hpp    LOM.cpp = src/lidar_odometry_mapping.cpp
#pragma once
#include <stdint.h>
#include <vector>

namespace vo {

struct P4 {  // pcl::PointXYZI restated as 16 bytes (common.h:42)
  float x, y, z, i;
};
typedef std::vector<P4> Cloud;

struct Params {
  int n_scans = 64;              // scan_line            (SR.cpp:50)
  double minimum_range = 5.0;    // minimum_range        (SR.cpp:53)
  float line_res = 0.4f;         // mapping_line_resolution  (LM.cpp:99)
  float plane_res = 0.8f;        // mapping_plane_resolution (LM.cpp:101)
  int mapping_skip_frame = 1;    // (LO.cpp:53, LM.cpp:125)
  int knn_backend = 0;           // 0 = brute force (truth), 1 = KD-tree (timing baseline)
  int distortion = 0;            // LaserOdometry::DISTORTION (LO.h:90, `false` in the reference): 1 = de-skew with s = relTime
};

// ---- third-party semantics -------------------------------------------------
void voxel_grid(const Cloud& in, float leaf, Cloud& out);  // pcl::VoxelGrid

struct KdTree;  // single KD-tree, leaf size 15, exact search, (d2, index) order
KdTree* kd_build(const Cloud& c);
void kd_free(KdTree* t);
// k nearest, sorted by (d2 f32, index); returns number found (min(k, n)).
int kd_knn(const KdTree* t, const Cloud& c, const P4& q, int k, int* idx, float* d2);
int brute_knn(const Cloud& c, const P4& q, int k, int* idx, float* d2);

void sym_eig3(const double A[9], double evals[3], double evecs[9]);  // ascending, columns
bool colpiv_qr_solve_5x3(const double A[15], const double b[5], double x[3]);
bool fit_line5(const float near[15], double a[3], double b[3]);     // LM.cpp:559-603 on five neighbours
bool fit_plane5(const float near[15], double nrm[3], double* d);    // LM.cpp:637-680

// ---- Ceres restatement -----------------------------------------------------
struct Factor {
  int type;     // 0 = LidarEdgeFactor, 1 = LidarPlaneFactor, 2 = LidarPlaneNormFactor
  double p[3];  // curr_point
  double a[3];  // edge: last_point_a | plane: last_point_j | planeNorm: unit normal
  double b[3];  // edge: last_point_b | plane: ljm_norm     | planeNorm: {d, -, -}
  double s = 1.0;        // interpolation ratio of LidarEdgeFactor / LidarPlaneFactor (LF.hpp:17, 66); 1.0 unless DISTORTION
  bool slerp = false;    // evaluate Identity.slerp(s, q) and s * t literally (LF.hpp:29-33, 86-90) instead of the s == 1 shortcut
};
struct SolveLog {
  int iterations = 0;       // attempted LM iterations (<= 4)
  int successful = 0;
  double initial_cost = 0, final_cost = 0;
  std::vector<double> cost_trace;  // cost after each attempt
};
void lm_register_full_cloud(Cloud& cloud, const double pose[7]);  // LM.cpp:901-905
// ceres::Solve as configured at LO.cpp:500-509 / LM.cpp:710-717.  x = {qx,qy,qz,qw,tx,ty,tz}.
void ceres_solve(const std::vector<Factor>& f, double x[7], SolveLog* log);
// One evaluation (robustified): cost, 6x6 J^T J (row-major), 6 J^T r.  For tests.
void evaluate_normal_eq(const std::vector<Factor>& f, const double x[7], double* cost, double H[36], double g[6]);

// ---- scanRegistration ------------------------------------------------------
struct ScanRegistration {
  Params prm;
  Cloud laserCloud, cornerSharp, cornerLessSharp, surfFlat, surfLessFlat;
  std::vector<float> curvature;
  std::vector<int> label, picked, sortInd;
  std::vector<int> scanStartInd, scanEndInd;
  float startOri = 0, endOri = 0;
  void reset();                                         // SR.cpp:95-104
  void input(const float* xyz, int n, int stride);      // SR.cpp:144-513
};

// ---- laserOdometry ---------------------------------------------------------
struct LaserOdometry {
  Params prm;
  Cloud cornerSharp, cornerLessSharp, surfFlat, surfLessFlat, fullRes;
  Cloud cornerLast, surfLast;
  KdTree *kdCorner = nullptr, *kdSurf = nullptr;
  double q_w[4] = {0, 0, 0, 1}, t_w[3] = {0, 0, 0};  // x,y,z,w
  double para_q[4] = {0, 0, 0, 1}, para_t[3] = {0, 0, 0};
  bool systemInited = false;
  int frameCount = 0;
  int corner_correspondence = 0, plane_correspondence = 0;
  // debug: association of the last solveLO, per outer pass
  std::vector<int> dbg_corner[2];  // 2 ints per sharp point (closest, second) or -1
  std::vector<int> dbg_surf[2];    // 3 ints per flat point
  SolveLog dbg_log[2];
  ~LaserOdometry();
  void input(const Cloud& full, const Cloud& sharp, const Cloud& lessSharp, const Cloud& flat,
             const Cloud& lessFlat);  // LO.cpp:137-148
  // association only, at pose (q,t); out arrays sized 2*nSharp / 3*nFlat
  void associate(const double q[4], const double t[3], std::vector<Factor>* f, std::vector<int>* ci,
                 std::vector<int>* si);
  void solveLO(const double* prior_q, const double* prior_t, bool use_prior);  // LO.cpp:199-584
  void set_last(const Cloud& corner, const Cloud& surf);
  bool skip_frame() const { return frameCount % prm.mapping_skip_frame != 0; }  // LO.cpp:668
};

// ---- laserMapping ----------------------------------------------------------
struct LaserMapping {
  enum { W = 21, H = 21, D = 11, NUM = W * H * D };  // LM.h:117-122
  Params prm;
  int cenW = 10, cenH = 10, cenD = 5;                // LM.h:75-78
  std::vector<Cloud*> cornerArray, surfArray;        // 4851 cubes each
  Cloud cornerLast, surfLast, cornerFromMap, surfFromMap, cornerStack, surfStack;
  int validInd[125];
  int validNum = 0;
  double parameters[7] = {0, 0, 0, 1, 0, 0, 0};      // q_w_curr (xyzw), t_w_curr
  double q_wmap_wodom[4] = {0, 0, 0, 1}, t_wmap_wodom[3] = {0, 0, 0};
  double q_wodom[4] = {0, 0, 0, 1}, t_wodom[3] = {0, 0, 0};
  double q_hf[4] = {0, 0, 0, 1}, t_hf[3] = {0, 0, 0};
  bool skip = false;
  int frameCount = 0;
  // debug of the last solveMapping, per outer pass: 5 idx + 5 d2 per query, accepted flag
  std::vector<int> dbg_cidx[2], dbg_sidx[2];
  std::vector<float> dbg_cd2[2], dbg_sd2[2];
  std::vector<int> dbg_cok[2], dbg_sok[2];
  SolveLog dbg_log[2];
  bool optimized = false;
  KdTree *kdCornerMap = nullptr, *kdSurfMap = nullptr;
  double t_tree_ms = 0, t_filter_ms = 0, t_assoc_ms = 0, t_solve_ms = 0;
  LaserMapping();
  ~LaserMapping();
  void reset() { validNum = 0; }  // LM.cpp:132-136
  void input(const Cloud& cornerLast_, const Cloud& surfLast_, const double q_wodom_[4],
             const double t_wodom_[3], bool skip_frame);  // LM.cpp:178-209
  void solveMapping();                                    // LM.cpp:212-814
  void associate(const double pose[7], std::vector<Factor>* f, int pass);  // LM.cpp:545-681
};

// ---- glue (LOM.cpp:65-176) -------------------------------------------------
struct Pipeline {
  ScanRegistration sr;
  LaserOdometry lo;
  LaserMapping lm;
  double ms_sr = 0, ms_lo = 0, ms_lm = 0;
  explicit Pipeline(const Params& p);
  void process(const float* xyz, int n, int stride);  // MAIN.cpp:143-144, 186-190
};

// quaternion helpers (Eigen semantics, x,y,z,w storage)
void q_mul(const double a[4], const double b[4], double out[4]);
void q_rot(const double q[4], const double v[3], double out[3]);
void q_inv(const double q[4], double out[4]);

}  // namespace vo
