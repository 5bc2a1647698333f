// oracle_stages.cpp -- TEST INFRASTRUCTURE (see vloam_oracle.h header).
// Statement-by-statement CPU restatement of ScanRegistration::input,
// LaserOdometry::solveLO and LaserMapping::solveMapping.  <math.h> (not <cmath>)
// is included on purpose: the reference TU sees the C float overloads, so
// atan / sqrt / atan2 on float arguments resolve to glibc atanf / sqrtf / atan2f
// (SURVEY.md section 7.2 item 1).  Build with -O2 -ffp-contract=off, no -march.
#include <math.h>
#include <float.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include "vloam_oracle.h"

namespace vo {

static inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ===========================================================================
// ScanRegistration
// ===========================================================================
void ScanRegistration::reset() {  // SR.cpp:95-104
  laserCloud.clear(); cornerSharp.clear(); cornerLessSharp.clear(); surfFlat.clear(); surfLessFlat.clear();
}

void ScanRegistration::input(const float* xyz, int n, int stride) {  // SR.cpp:144-513
  const int N_SCANS = prm.n_scans;
  const double scanPeriod = 0.1;  // SR.h:84
  scanStartInd.assign(N_SCANS, 0);
  scanEndInd.assign(N_SCANS, 0);
  curvature.clear(); label.clear(); picked.clear(); sortInd.clear();

  // SR.cpp:174 removeNaNFromPointCloud (restated as an unconditional finite
  // check) + SR.cpp:107-141 removeClosedPointCloud, thres = (float)MINIMUM_RANGE.
  const float thres = (float)prm.minimum_range;
  std::vector<float> in; in.reserve((size_t)n * 3);
  for (int i = 0; i < n; ++i) {
    const float x = xyz[(size_t)i * stride], y = xyz[(size_t)i * stride + 1], z = xyz[(size_t)i * stride + 2];
    if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
    if (x * x + y * y + z * z < thres * thres) continue;
    in.push_back(x); in.push_back(y); in.push_back(z);
  }
  int cloudSize = (int)(in.size() / 3);
  if (cloudSize == 0) {  // the reference would index points[0]; defined here as "no output"
    for (int i = 0; i < N_SCANS; i++) { scanStartInd[i] = 5; scanEndInd[i] = -6; }  // what SR.cpp:308-315 yields for empty rings
    return;
  }

  // SR.cpp:183-197
  startOri = -atan2f(in[1], in[0]);
  endOri = -atan2f(in[(size_t)(cloudSize - 1) * 3 + 1], in[(size_t)(cloudSize - 1) * 3]) + 2 * M_PI;
  if (endOri - startOri > 3 * M_PI) endOri -= 2 * M_PI;
  else if (endOri - startOri < M_PI) endOri += 2 * M_PI;

  bool halfPassed = false;
  int count = cloudSize;
  std::vector<Cloud> laserCloudScans(N_SCANS);
  for (int i = 0; i < cloudSize; i++) {  // SR.cpp:209-298
    P4 point;
    point.x = in[(size_t)i * 3]; point.y = in[(size_t)i * 3 + 1]; point.z = in[(size_t)i * 3 + 2];
    float angle = atanf(point.z / sqrtf(point.x * point.x + point.y * point.y)) * 180 / M_PI;
    int scanID = 0;
    if (N_SCANS == 16) {
      scanID = int((angle + 15) / 2 + 0.5);
      if (scanID > (N_SCANS - 1) || scanID < 0) { count--; continue; }
    } else if (N_SCANS == 32) {
      scanID = int((angle + 92.0 / 3.0) * 3.0 / 4.0);
      if (scanID > (N_SCANS - 1) || scanID < 0) { count--; continue; }
    } else if (N_SCANS == 64) {
      if (angle >= -8.83) scanID = int((2 - angle) * 3.0 + 0.5);
      else scanID = N_SCANS / 2 + int((-8.83 - angle) * 2.0 + 0.5);
      if (angle > 2 || angle < -24.33 || scanID > 50 || scanID < 0) { count--; continue; }
    } else if (N_SCANS == 128) {
      // BUILDER EXTENSION: the reference aborts for
      // anything but 16/32/64.  128 uniform beams over +-22.5 deg, defined identically
      // in the CUDA path.
      const double v = (angle + 22.5) / (45.0 / 127.0) + 0.5;
      scanID = int(v);
      if (v < 0.0 || scanID > (N_SCANS - 1)) { count--; continue; }
    } else {
      return;  // ROS_BREAK in the reference
    }
    float ori = -atan2f(point.y, point.x);
    if (!halfPassed) {
      if (ori < startOri - M_PI / 2) ori += 2 * M_PI;
      else if (ori > startOri + M_PI * 3 / 2) ori -= 2 * M_PI;
      if (ori - startOri > M_PI) halfPassed = true;
    } else {
      ori += 2 * M_PI;
      if (ori < endOri - M_PI * 3 / 2) ori += 2 * M_PI;
      else if (ori > endOri + M_PI / 2) ori -= 2 * M_PI;
    }
    float relTime = (ori - startOri) / (endOri - startOri);
    point.i = scanID + scanPeriod * relTime;
    laserCloudScans[scanID].push_back(point);
  }
  cloudSize = count;

  for (int i = 0; i < N_SCANS; i++) {  // SR.cpp:308-315
    scanStartInd[i] = (int)laserCloud.size() + 5;
    laserCloud.insert(laserCloud.end(), laserCloudScans[i].begin(), laserCloudScans[i].end());
    scanEndInd[i] = (int)laserCloud.size() - 6;
  }

  curvature.assign(std::max(cloudSize, 0), 0.f);
  sortInd.assign(std::max(cloudSize, 0), 0);
  picked.assign(std::max(cloudSize, 0), 0);
  label.assign(std::max(cloudSize, 0), 0);
  const Cloud& lc = laserCloud;
  for (int i = 5; i < cloudSize - 5; i++) {  // SR.cpp:323-346
    float diffX = lc[i - 5].x + lc[i - 4].x + lc[i - 3].x + lc[i - 2].x + lc[i - 1].x - 10 * lc[i].x +
                  lc[i + 1].x + lc[i + 2].x + lc[i + 3].x + lc[i + 4].x + lc[i + 5].x;
    float diffY = lc[i - 5].y + lc[i - 4].y + lc[i - 3].y + lc[i - 2].y + lc[i - 1].y - 10 * lc[i].y +
                  lc[i + 1].y + lc[i + 2].y + lc[i + 3].y + lc[i + 4].y + lc[i + 5].y;
    float diffZ = lc[i - 5].z + lc[i - 4].z + lc[i - 3].z + lc[i - 2].z + lc[i - 1].z - 10 * lc[i].z +
                  lc[i + 1].z + lc[i + 2].z + lc[i + 3].z + lc[i + 4].z + lc[i + 5].z;
    curvature[i] = diffX * diffX + diffY * diffY + diffZ * diffZ;
    sortInd[i] = i;
    picked[i] = 0;
    label[i] = 0;
  }

  auto gap2 = [&](int a, int b) {
    float dX = lc[a].x - lc[b].x, dY = lc[a].y - lc[b].y, dZ = lc[a].z - lc[b].z;
    return dX * dX + dY * dY + dZ * dZ;
  };

  for (int i = 0; i < N_SCANS; i++) {  // SR.cpp:352-504
    if (scanEndInd[i] - scanStartInd[i] < 6) continue;
    Cloud lessFlatScan;
    for (int j = 0; j < 6; j++) {
      int sp = scanStartInd[i] + (scanEndInd[i] - scanStartInd[i]) * j / 6;
      int ep = scanStartInd[i] + (scanEndInd[i] - scanStartInd[i]) * (j + 1) / 6 - 1;
      // SR.cpp:365-366; canonical tie refinement: (curvature, index)
      std::sort(sortInd.begin() + sp, sortInd.begin() + ep + 1, [&](int a, int b) {
        return curvature[a] != curvature[b] ? curvature[a] < curvature[b] : a < b;
      });
      int largestPickedNum = 0;
      for (int k = ep; k >= sp; k--) {  // SR.cpp:371-431
        int ind = sortInd[k];
        if (picked[ind] == 0 && curvature[ind] > 0.1) {
          largestPickedNum++;
          if (largestPickedNum <= 2) {
            label[ind] = 2;
            cornerSharp.push_back(lc[ind]);
            cornerLessSharp.push_back(lc[ind]);
          } else if (largestPickedNum <= 20) {
            label[ind] = 1;
            cornerLessSharp.push_back(lc[ind]);
          } else break;
          picked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            if (gap2(ind + l, ind + l - 1) > 0.05) break;
            picked[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            if (gap2(ind + l, ind + l + 1) > 0.05) break;
            picked[ind + l] = 1;
          }
        }
      }
      int smallestPickedNum = 0;
      for (int k = sp; k <= ep; k++) {  // SR.cpp:439-483
        int ind = sortInd[k];
        if (picked[ind] == 0 && curvature[ind] < 0.1) {
          label[ind] = -1;
          surfFlat.push_back(lc[ind]);
          smallestPickedNum++;
          if (smallestPickedNum >= 4) break;
          picked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            if (gap2(ind + l, ind + l - 1) > 0.05) break;
            picked[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            if (gap2(ind + l, ind + l + 1) > 0.05) break;
            picked[ind + l] = 1;
          }
        }
      }
      for (int k = sp; k <= ep; k++)  // SR.cpp:486-493
        if (label[k] <= 0) lessFlatScan.push_back(lc[k]);
    }
    Cloud ds;
    voxel_grid(lessFlatScan, 0.2f, ds);  // SR.cpp:497-501 (setLeafSize takes floats)
    surfLessFlat.insert(surfLessFlat.end(), ds.begin(), ds.end());
  }
}

// ===========================================================================
// LaserOdometry
// ===========================================================================
LaserOdometry::~LaserOdometry() { if (kdCorner) kd_free(kdCorner); if (kdSurf) kd_free(kdSurf); }

void LaserOdometry::input(const Cloud& full, const Cloud& sharp, const Cloud& lessSharp, const Cloud& flat,
                          const Cloud& lessFlat) {  // LO.cpp:137-148 (deep copies)
  fullRes = full; cornerSharp = sharp; cornerLessSharp = lessSharp; surfFlat = flat; surfLessFlat = lessFlat;
}

void LaserOdometry::set_last(const Cloud& corner, const Cloud& surf) {
  cornerLast = corner; surfLast = surf;
  if (kdCorner) kd_free(kdCorner); if (kdSurf) kd_free(kdSurf);
  kdCorner = nullptr; kdSurf = nullptr;
  if (prm.knn_backend == 1) { kdCorner = kd_build(cornerLast); kdSurf = kd_build(surfLast); }
  systemInited = true;
}

void slerp_identity_d(double t, const double q[4], double out[4]);  // oracle_ceres.cpp: Eigen's Identity.slerp(t, q)

// interpolation ratio of a point (LO.cpp:156-160, 371, 475): (intensity - int(intensity)) is a float, SCAN_PERIOD a double
static inline double point_s(const P4& pi, bool distortion) {
  const double SCAN_PERIOD = 0.1;  // LO.h:93
  if (distortion) return (pi.i - int(pi.i)) / SCAN_PERIOD;
  return 1.0;
}

// TransformToStart (LO.cpp:152-173).  DISTORTION == false: s = 1 and Identity.slerp(1, q) == +-q;
// DISTORTION == true: q_point_last = Identity.slerp(s, q), t_point_last = s * t.  f64 math stored to f32.
static inline P4 transform_to_start(const P4& pi, const double q[4], const double t[3], bool distortion) {
  const double v[3] = {pi.x, pi.y, pi.z};
  const double s = point_s(pi, distortion);
  double qs[4] = {q[0], q[1], q[2], q[3]};
  if (distortion) slerp_identity_d(s, q, qs);
  double r[3]; q_rot(qs, v, r);
  const double tl[3] = {s * t[0], s * t[1], s * t[2]};
  P4 po; po.x = (float)(r[0] + tl[0]); po.y = (float)(r[1] + tl[1]); po.z = (float)(r[2] + tl[2]); po.i = pi.i;
  return po;
}

static inline double sqdis(const P4& a, const P4& s) {  // LO.cpp:319-322: float ops widened to double
  return (a.x - s.x) * (a.x - s.x) + (a.y - s.y) * (a.y - s.y) + (a.z - s.z) * (a.z - s.z);
}

void LaserOdometry::associate(const double q[4], const double t[3], std::vector<Factor>* fs,
                              std::vector<int>* ci, std::vector<int>* si) {
  const double DISTANCE_SQ_THRESHOLD = 25, NEARBY_SCAN = 2.5;  // LO.h:94-95
  const int nSharp = (int)cornerSharp.size(), nFlat = (int)surfFlat.size();
  if (ci) ci->assign((size_t)nSharp * 2, -1);
  if (si) si->assign((size_t)nFlat * 3, -1);
  corner_correspondence = 0; plane_correspondence = 0;
  const Cloud& CL = cornerLast; const Cloud& SL = surfLast;
  for (int i = 0; i < nSharp; ++i) {  // LO.cpp:282-383
    const P4 pointSel = transform_to_start(cornerSharp[i], q, t, prm.distortion != 0);
    int nn; float nd;
    const int found = kdCorner ? kd_knn(kdCorner, CL, pointSel, 1, &nn, &nd) : brute_knn(CL, pointSel, 1, &nn, &nd);
    int closestPointInd = -1, minPointInd2 = -1;
    if (found > 0 && nd < DISTANCE_SQ_THRESHOLD) {
      closestPointInd = nn;
      int closestPointScanID = int(CL[closestPointInd].i);
      double minPointSqDis2 = DISTANCE_SQ_THRESHOLD;
      for (int j = closestPointInd + 1; j < (int)CL.size(); ++j) {
        if (int(CL[j].i) <= closestPointScanID) continue;
        if (int(CL[j].i) > (closestPointScanID + NEARBY_SCAN)) break;
        double pointSqDis = sqdis(CL[j], pointSel);
        if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
      }
      for (int j = closestPointInd - 1; j >= 0; --j) {
        if (int(CL[j].i) >= closestPointScanID) continue;
        if (int(CL[j].i) < (closestPointScanID - NEARBY_SCAN)) break;
        double pointSqDis = sqdis(CL[j], pointSel);
        if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
      }
    }
    if (ci) { (*ci)[(size_t)i * 2] = closestPointInd; (*ci)[(size_t)i * 2 + 1] = minPointInd2; }
    if (minPointInd2 >= 0) {
      Factor f; f.type = 0;
      f.p[0] = cornerSharp[i].x; f.p[1] = cornerSharp[i].y; f.p[2] = cornerSharp[i].z;
      f.a[0] = CL[closestPointInd].x; f.a[1] = CL[closestPointInd].y; f.a[2] = CL[closestPointInd].z;
      f.b[0] = CL[minPointInd2].x; f.b[1] = CL[minPointInd2].y; f.b[2] = CL[minPointInd2].z;
      f.s = point_s(cornerSharp[i], prm.distortion != 0); f.slerp = prm.distortion != 0;  // LO.cpp:368-372
      if (fs) fs->push_back(f);
      corner_correspondence++;
    }
  }
  for (int i = 0; i < nFlat; ++i) {  // LO.cpp:387-485
    const P4 pointSel = transform_to_start(surfFlat[i], q, t, prm.distortion != 0);
    int nn; float nd;
    const int found = kdSurf ? kd_knn(kdSurf, SL, pointSel, 1, &nn, &nd) : brute_knn(SL, pointSel, 1, &nn, &nd);
    int closestPointInd = -1, minPointInd2 = -1, minPointInd3 = -1;
    if (found > 0 && nd < DISTANCE_SQ_THRESHOLD) {
      closestPointInd = nn;
      int closestPointScanID = int(SL[closestPointInd].i);
      double minPointSqDis2 = DISTANCE_SQ_THRESHOLD, minPointSqDis3 = DISTANCE_SQ_THRESHOLD;
      for (int j = closestPointInd + 1; j < (int)SL.size(); ++j) {
        if (int(SL[j].i) > (closestPointScanID + NEARBY_SCAN)) break;
        double pointSqDis = sqdis(SL[j], pointSel);
        if (int(SL[j].i) <= closestPointScanID && pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
        else if (int(SL[j].i) > closestPointScanID && pointSqDis < minPointSqDis3) { minPointSqDis3 = pointSqDis; minPointInd3 = j; }
      }
      for (int j = closestPointInd - 1; j >= 0; --j) {
        if (int(SL[j].i) < (closestPointScanID - NEARBY_SCAN)) break;
        double pointSqDis = sqdis(SL[j], pointSel);
        if (int(SL[j].i) >= closestPointScanID && pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
        else if (int(SL[j].i) < closestPointScanID && pointSqDis < minPointSqDis3) { minPointSqDis3 = pointSqDis; minPointInd3 = j; }
      }
    }
    if (si) { (*si)[(size_t)i * 3] = closestPointInd; (*si)[(size_t)i * 3 + 1] = minPointInd2; (*si)[(size_t)i * 3 + 2] = minPointInd3; }
    if (closestPointInd >= 0 && minPointInd2 >= 0 && minPointInd3 >= 0) {
      Factor f; f.type = 1;
      f.p[0] = surfFlat[i].x; f.p[1] = surfFlat[i].y; f.p[2] = surfFlat[i].z;
      const double J[3] = {SL[closestPointInd].x, SL[closestPointInd].y, SL[closestPointInd].z};
      const double L[3] = {SL[minPointInd2].x, SL[minPointInd2].y, SL[minPointInd2].z};
      const double M[3] = {SL[minPointInd3].x, SL[minPointInd3].y, SL[minPointInd3].z};
      // LidarPlaneFactor ctor LF.hpp:73-74: ljm_norm = normalize((j-l) x (j-m))
      const double u[3] = {J[0] - L[0], J[1] - L[1], J[2] - L[2]}, w[3] = {J[0] - M[0], J[1] - M[1], J[2] - M[2]};
      double nrm[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
      const double n2 = nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2];
      if (n2 > 0) { const double n = sqrt(n2); nrm[0] /= n; nrm[1] /= n; nrm[2] /= n; }  // Eigen normalize(): only if z > 0
      for (int k = 0; k < 3; ++k) { f.a[k] = J[k]; f.b[k] = nrm[k]; }
      f.s = point_s(surfFlat[i], prm.distortion != 0); f.slerp = prm.distortion != 0;  // LO.cpp:472-476
      if (fs) fs->push_back(f);
      plane_correspondence++;
    }
  }
}

void LaserOdometry::solveLO(const double* prior_q, const double* prior_t, bool use_prior) {  // LO.cpp:199-584
  for (int k = 0; k < 2; ++k) { dbg_corner[k].clear(); dbg_surf[k].clear(); dbg_log[k] = SolveLog(); }
  if (!systemInited) {
    systemInited = true;
  } else {
    for (size_t opti_counter = 0; opti_counter < 2; ++opti_counter) {
      if (use_prior) {  // LO.cpp:237-250 (detach_VO_LO == false)
        for (int k = 0; k < 4; ++k) para_q[k] = prior_q[k];
        for (int k = 0; k < 3; ++k) para_t[k] = prior_t[k];
      }
      std::vector<Factor> fs;
      associate(para_q, para_t, &fs, &dbg_corner[opti_counter], &dbg_surf[opti_counter]);
      double x[7] = {para_q[0], para_q[1], para_q[2], para_q[3], para_t[0], para_t[1], para_t[2]};
      ceres_solve(fs, x, &dbg_log[opti_counter]);
      for (int k = 0; k < 4; ++k) para_q[k] = x[k];
      for (int k = 0; k < 3; ++k) para_t[k] = x[4 + k];
    }
    // LO.cpp:524-525
    double r[3]; q_rot(q_w, para_t, r);
    t_w[0] = t_w[0] + r[0]; t_w[1] = t_w[1] + r[1]; t_w[2] = t_w[2] + r[2];
    double qn[4]; q_mul(q_w, para_q, qn);
    for (int k = 0; k < 4; ++k) q_w[k] = qn[k];
  }
  // LO.cpp:558-574: swap in the new "last" clouds, rebuild both KD-trees
  cornerLessSharp.swap(cornerLast);
  surfLessFlat.swap(surfLast);
  if (kdCorner) { kd_free(kdCorner); kdCorner = nullptr; }
  if (kdSurf) { kd_free(kdSurf); kdSurf = nullptr; }
  if (prm.knn_backend == 1) { kdCorner = kd_build(cornerLast); kdSurf = kd_build(surfLast); }
  frameCount++;
}

// ===========================================================================
// LaserMapping
// ===========================================================================
LaserMapping::LaserMapping() : cornerArray(NUM), surfArray(NUM) {
  for (int i = 0; i < NUM; ++i) { cornerArray[i] = new Cloud(); surfArray[i] = new Cloud(); }
}
LaserMapping::~LaserMapping() { for (int i = 0; i < NUM; ++i) { delete cornerArray[i]; delete surfArray[i]; } }

void LaserMapping::input(const Cloud& cornerLast_, const Cloud& surfLast_, const double q_wodom_[4],
                         const double t_wodom_[3], bool skip_frame) {  // LM.cpp:178-209
  skip = skip_frame;
  if (!skip) { cornerLast = cornerLast_; surfLast = surfLast_; }
  for (int k = 0; k < 4; ++k) q_wodom[k] = q_wodom_[k];
  for (int k = 0; k < 3; ++k) t_wodom[k] = t_wodom_[k];
  double r[3]; q_rot(q_wmap_wodom, t_wodom, r);
  if (skip) {
    q_mul(q_wmap_wodom, q_wodom, q_hf);
    for (int k = 0; k < 3; ++k) t_hf[k] = r[k] + t_wmap_wodom[k];
  } else {
    double qn[4]; q_mul(q_wmap_wodom, q_wodom, qn);
    for (int k = 0; k < 4; ++k) parameters[k] = qn[k];
    for (int k = 0; k < 3; ++k) parameters[4 + k] = r[k] + t_wmap_wodom[k];
  }
}

static inline P4 associate_to_map(const P4& pi, const double pose[7]) {  // LM.cpp:154-164
  const double v[3] = {pi.x, pi.y, pi.z};
  double r[3]; q_rot(pose, v, r);
  P4 po; po.x = (float)(r[0] + pose[4]); po.y = (float)(r[1] + pose[5]); po.z = (float)(r[2] + pose[6]); po.i = pi.i;
  return po;
}

// LM.cpp:901-905: every point of laserCloudFullRes through pointAssociateToMap, in place (LaserMapping::publish)
void lm_register_full_cloud(Cloud& cloud, const double pose[7]) {
  for (auto& pt : cloud) pt = associate_to_map(pt, pose);
}

void LaserMapping::associate(const double pose[7], std::vector<Factor>* fs, int pass) {  // LM.cpp:545-681
  KdTree* kc = kdCornerMap; KdTree* ks = kdSurfMap;
  const int nc = (int)cornerStack.size(), ns = (int)surfStack.size();
  std::vector<int>& cidx = dbg_cidx[pass]; std::vector<int>& sidx = dbg_sidx[pass];
  std::vector<float>& cd2 = dbg_cd2[pass]; std::vector<float>& sd2 = dbg_sd2[pass];
  cidx.assign((size_t)nc * 5, -1); cd2.assign((size_t)nc * 5, 0.f); dbg_cok[pass].assign(nc, 0);
  sidx.assign((size_t)ns * 5, -1); sd2.assign((size_t)ns * 5, 0.f); dbg_sok[pass].assign(ns, 0);
  for (int i = 0; i < nc; i++) {
    const P4 pointOri = cornerStack[i];
    const P4 pointSel = associate_to_map(pointOri, pose);
    int ind[5]; float d2[5];
    const int found = kc ? kd_knn(kc, cornerFromMap, pointSel, 5, ind, d2) : brute_knn(cornerFromMap, pointSel, 5, ind, d2);
    if (found < 5) continue;
    for (int j = 0; j < 5; ++j) { cidx[(size_t)i * 5 + j] = ind[j]; cd2[(size_t)i * 5 + j] = d2[j]; }
    if (d2[4] < 1.0) {
      float near[15];
      for (int j = 0; j < 5; j++) { near[j * 3] = cornerFromMap[ind[j]].x; near[j * 3 + 1] = cornerFromMap[ind[j]].y; near[j * 3 + 2] = cornerFromMap[ind[j]].z; }
      Factor f; f.type = 0;
      if (fit_line5(near, f.a, f.b)) {  // LM.cpp:559-603
        f.p[0] = pointOri.x; f.p[1] = pointOri.y; f.p[2] = pointOri.z;
        if (fs) fs->push_back(f);
        dbg_cok[pass][i] = 1;
      }
    }
  }
  for (int i = 0; i < ns; i++) {
    const P4 pointOri = surfStack[i];
    const P4 pointSel = associate_to_map(pointOri, pose);
    int ind[5]; float d2[5];
    const int found = ks ? kd_knn(ks, surfFromMap, pointSel, 5, ind, d2) : brute_knn(surfFromMap, pointSel, 5, ind, d2);
    if (found < 5) continue;
    for (int j = 0; j < 5; ++j) { sidx[(size_t)i * 5 + j] = ind[j]; sd2[(size_t)i * 5 + j] = d2[j]; }
    if (d2[4] < 1.0) {
      float near[15];
      for (int j = 0; j < 5; j++) { near[j * 3] = surfFromMap[ind[j]].x; near[j * 3 + 1] = surfFromMap[ind[j]].y; near[j * 3 + 2] = surfFromMap[ind[j]].z; }
      Factor f; f.type = 2;
      double d = 0;
      if (fit_plane5(near, f.a, &d)) {  // LM.cpp:637-680
        f.p[0] = pointOri.x; f.p[1] = pointOri.y; f.p[2] = pointOri.z;
        f.b[0] = d; f.b[1] = 0; f.b[2] = 0;
        if (fs) fs->push_back(f);
        dbg_sok[pass][i] = 1;
      }
    }
  }
}

void LaserMapping::solveMapping() {  // LM.cpp:212-814
  double* t = parameters + 4;
  int centerCubeI = int((t[0] + 25.0) / 50.0) + cenW;  // LM.cpp:228-241
  int centerCubeJ = int((t[1] + 25.0) / 50.0) + cenH;
  int centerCubeK = int((t[2] + 25.0) / 50.0) + cenD;
  if (t[0] + 25.0 < 0) centerCubeI--;
  if (t[1] + 25.0 < 0) centerCubeJ--;
  if (t[2] + 25.0 < 0) centerCubeK--;
  auto IDX = [](int i, int j, int k) { return i + W * j + W * H * k; };
  // LM.cpp:252-444: six roll loops; the slab that wraps is cleared.
  while (centerCubeI < 3) {
    for (int j = 0; j < H; j++) for (int k = 0; k < D; k++) {
      int i = W - 1; Cloud* pc = cornerArray[IDX(i, j, k)]; Cloud* ps = surfArray[IDX(i, j, k)];
      for (; i >= 1; i--) { cornerArray[IDX(i, j, k)] = cornerArray[IDX(i - 1, j, k)]; surfArray[IDX(i, j, k)] = surfArray[IDX(i - 1, j, k)]; }
      cornerArray[IDX(i, j, k)] = pc; surfArray[IDX(i, j, k)] = ps; pc->clear(); ps->clear();
    }
    centerCubeI++; cenW++;
  }
  while (centerCubeI >= W - 3) {
    for (int j = 0; j < H; j++) for (int k = 0; k < D; k++) {
      int i = 0; Cloud* pc = cornerArray[IDX(i, j, k)]; Cloud* ps = surfArray[IDX(i, j, k)];
      for (; i < W - 1; i++) { cornerArray[IDX(i, j, k)] = cornerArray[IDX(i + 1, j, k)]; surfArray[IDX(i, j, k)] = surfArray[IDX(i + 1, j, k)]; }
      cornerArray[IDX(i, j, k)] = pc; surfArray[IDX(i, j, k)] = ps; pc->clear(); ps->clear();
    }
    centerCubeI--; cenW--;
  }
  while (centerCubeJ < 3) {
    for (int i = 0; i < W; i++) for (int k = 0; k < D; k++) {
      int j = H - 1; Cloud* pc = cornerArray[IDX(i, j, k)]; Cloud* ps = surfArray[IDX(i, j, k)];
      for (; j >= 1; j--) { cornerArray[IDX(i, j, k)] = cornerArray[IDX(i, j - 1, k)]; surfArray[IDX(i, j, k)] = surfArray[IDX(i, j - 1, k)]; }
      cornerArray[IDX(i, j, k)] = pc; surfArray[IDX(i, j, k)] = ps; pc->clear(); ps->clear();
    }
    centerCubeJ++; cenH++;
  }
  while (centerCubeJ >= H - 3) {
    for (int i = 0; i < W; i++) for (int k = 0; k < D; k++) {
      int j = 0; Cloud* pc = cornerArray[IDX(i, j, k)]; Cloud* ps = surfArray[IDX(i, j, k)];
      for (; j < H - 1; j++) { cornerArray[IDX(i, j, k)] = cornerArray[IDX(i, j + 1, k)]; surfArray[IDX(i, j, k)] = surfArray[IDX(i, j + 1, k)]; }
      cornerArray[IDX(i, j, k)] = pc; surfArray[IDX(i, j, k)] = ps; pc->clear(); ps->clear();
    }
    centerCubeJ--; cenH--;
  }
  while (centerCubeK < 3) {
    for (int i = 0; i < W; i++) for (int j = 0; j < H; j++) {
      int k = D - 1; Cloud* pc = cornerArray[IDX(i, j, k)]; Cloud* ps = surfArray[IDX(i, j, k)];
      for (; k >= 1; k--) { cornerArray[IDX(i, j, k)] = cornerArray[IDX(i, j, k - 1)]; surfArray[IDX(i, j, k)] = surfArray[IDX(i, j, k - 1)]; }
      cornerArray[IDX(i, j, k)] = pc; surfArray[IDX(i, j, k)] = ps; pc->clear(); ps->clear();
    }
    centerCubeK++; cenD++;
  }
  while (centerCubeK >= D - 3) {
    for (int i = 0; i < W; i++) for (int j = 0; j < H; j++) {
      int k = 0; Cloud* pc = cornerArray[IDX(i, j, k)]; Cloud* ps = surfArray[IDX(i, j, k)];
      for (; k < D - 1; k++) { cornerArray[IDX(i, j, k)] = cornerArray[IDX(i, j, k + 1)]; surfArray[IDX(i, j, k)] = surfArray[IDX(i, j, k + 1)]; }
      cornerArray[IDX(i, j, k)] = pc; surfArray[IDX(i, j, k)] = ps; pc->clear(); ps->clear();
    }
    centerCubeK--; cenD--;
  }
  // LM.cpp:448-466
  for (int i = centerCubeI - 2; i <= centerCubeI + 2; i++)
    for (int j = centerCubeJ - 2; j <= centerCubeJ + 2; j++)
      for (int k = centerCubeK - 1; k <= centerCubeK + 1; k++)
        if (i >= 0 && i < W && j >= 0 && j < H && k >= 0 && k < D) validInd[validNum++] = IDX(i, j, k);
  // LM.cpp:476-485
  cornerFromMap.clear(); surfFromMap.clear();
  for (int i = 0; i < validNum; i++) {
    const Cloud& c = *cornerArray[validInd[i]]; const Cloud& s = *surfArray[validInd[i]];
    cornerFromMap.insert(cornerFromMap.end(), c.begin(), c.end());
    surfFromMap.insert(surfFromMap.end(), s.begin(), s.end());
  }
  // LM.cpp:492-500
  voxel_grid(cornerLast, prm.line_res, cornerStack);
  voxel_grid(surfLast, prm.plane_res, surfStack);

  for (int p = 0; p < 2; ++p) { dbg_cidx[p].clear(); dbg_sidx[p].clear(); dbg_cd2[p].clear(); dbg_sd2[p].clear(); dbg_cok[p].clear(); dbg_sok[p].clear(); dbg_log[p] = SolveLog(); }
  optimized = false;
  t_tree_ms = t_assoc_ms = t_solve_ms = t_filter_ms = 0;
  if ((int)cornerFromMap.size() > 10 && (int)surfFromMap.size() > 50) {  // LM.cpp:514
    optimized = true;
    if (prm.knn_backend == 1) {  // LM.cpp:519-520 setInputCloud = full KD-tree rebuild
      const double t0 = now_ms();
      kdCornerMap = kd_build(cornerFromMap); kdSurfMap = kd_build(surfFromMap);
      t_tree_ms = now_ms() - t0;
    }
    for (int iterCount = 0; iterCount < 2; iterCount++) {  // LM.cpp:526
      std::vector<Factor> fs;
      double t0 = now_ms();
      associate(parameters, &fs, iterCount);
      t_assoc_ms += now_ms() - t0; t0 = now_ms();
      ceres_solve(fs, parameters, &dbg_log[iterCount]);
      t_solve_ms += now_ms() - t0;
    }
    if (kdCornerMap) { kd_free(kdCornerMap); kdCornerMap = nullptr; }
    if (kdSurfMap) { kd_free(kdSurfMap); kdSurfMap = nullptr; }
  }
  // transformUpdate LM.cpp:147-151, 737
  {
    double qi[4]; q_inv(q_wodom, qi);
    q_mul(parameters, qi, q_wmap_wodom);
    double r[3]; q_rot(q_wmap_wodom, t_wodom, r);
    for (int k = 0; k < 3; ++k) t_wmap_wodom[k] = parameters[4 + k] - r[k];
  }
  // LM.cpp:741-788
  auto insert = [&](const Cloud& stack, std::vector<Cloud*>& arr) {
    for (const P4& p : stack) {
      const P4 pointSel = associate_to_map(p, parameters);
      int cubeI = int((pointSel.x + 25.0) / 50.0) + cenW;
      int cubeJ = int((pointSel.y + 25.0) / 50.0) + cenH;
      int cubeK = int((pointSel.z + 25.0) / 50.0) + cenD;
      if (pointSel.x + 25.0 < 0) cubeI--;
      if (pointSel.y + 25.0 < 0) cubeJ--;
      if (pointSel.z + 25.0 < 0) cubeK--;
      if (cubeI >= 0 && cubeI < W && cubeJ >= 0 && cubeJ < H && cubeK >= 0 && cubeK < D)
        arr[IDX(cubeI, cubeJ, cubeK)]->push_back(pointSel);
    }
  };
  insert(cornerStack, cornerArray);
  insert(surfStack, surfArray);
  // LM.cpp:795-808
  const double t0 = now_ms();
  for (int i = 0; i < validNum; i++) {
    const int ind = validInd[i];
    Cloud tmp;
    voxel_grid(*cornerArray[ind], prm.line_res, tmp); cornerArray[ind]->swap(tmp);
    voxel_grid(*surfArray[ind], prm.plane_res, tmp); surfArray[ind]->swap(tmp);
  }
  t_filter_ms = now_ms() - t0;
  frameCount++;
}

// ===========================================================================
// glue: MAIN.cpp:143-144, 186-190 -> LOM.cpp:65-176
// ===========================================================================
Pipeline::Pipeline(const Params& p) { sr.prm = p; lo.prm = p; lm.prm = p; }

void Pipeline::process(const float* xyz, int n, int stride) {
  sr.reset(); lm.reset();
  double t0 = now_ms();
  sr.input(xyz, n, stride);
  ms_sr = now_ms() - t0; t0 = now_ms();
  lo.input(sr.laserCloud, sr.cornerSharp, sr.cornerLessSharp, sr.surfFlat, sr.surfLessFlat);
  lo.solveLO(nullptr, nullptr, false);
  const bool skip = lo.skip_frame();
  ms_lo = now_ms() - t0; t0 = now_ms();
  lm.input(lo.cornerLast, lo.surfLast, lo.q_w, lo.t_w, skip);
  if (!skip) lm.solveMapping();
  ms_lm = now_ms() - t0;
}

}  // namespace vo
