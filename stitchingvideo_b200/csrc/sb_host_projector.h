// sb_host_projector.h — host-side projector arithmetic (per calibration, O(perimeter)).
#pragma once
#include "sb_kernels.h"

namespace sb {
// ProjectorBase::setCameraParams (warpers.cpp:50-78)
void projector_set(ProjParams &p, int kind, float scale, const float K[9], const float R[9], const float T[3]);
// {Plane,Spherical,Cylindrical}Projector::mapForward (warpers_inl.hpp:206-218,237-247,271-280)
void projector_map_forward(const ProjParams &p, float x, float y, float *u, float *v);
// detectResultRoi: PlaneWarper (warpers.cpp:139-168), by-border (warpers_inl.hpp:169-203),
// SphericalWarper pole handling (warpers.cpp:171-212).  tl/br are inclusive corners.
void projector_detect_result_roi(const ProjParams &p, int src_w, int src_h, sb_point *tl, sb_point *br);
}  // namespace sb
