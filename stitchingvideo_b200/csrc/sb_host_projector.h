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
// mapBackward of every projector on the host (warpers_inl.hpp:222-759): used for the kinds whose maps are host-built
void projector_map_backward(const ProjParams &p, float u, float v, float *x, float *y);
// buildMaps (warpers_inl.hpp:62-85) into host arrays of (br.y - tl.y + 1) x (br.x - tl.x + 1) floats
void projector_build_maps_host(const ProjParams &p, sb_point tl, sb_point br, float *xmap, float *ymap);
void projector_detect_result_roi(const ProjParams &p, int src_w, int src_h, sb_point *tl, sb_point *br);
}  // namespace sb
