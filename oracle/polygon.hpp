// TEST INFRASTRUCTURE - CPU restatement of the plane / polygon matching step that follows find_primitives:
// MapPlane::find_matches (src/map_management/map_features/map_primitive.cpp:91-161), WorldPolygon::to_camera_space
// (src/coordinates/polygon_coordinates.cpp:135-162), Polygon::project / inter_area (src/utils/polygon.cpp:349-382,542-561),
// Plane::is_distance_similar / is_normal_similar (src/features/primitives/shape_primitives.cpp:66-84).
// boost::geometry (intersection, area, correct) is an un-vendored third-party dependency of the reference: its published
// semantics are restated - the summed area of the intersection of two valid simple polygons - by a vertical-slab sweep,
// pinned in tests/test_oracle_polygon.py against closed forms, the real OpenCV (cv2.intersectConvexConvex) and rasterisation.
// PARITY UNPINNED against boost itself (absent from this machine).
#pragma once
#include "../include/rgbdslam_b200.h"

namespace oracle {

// area of the intersection of two simple polygons given as open rings of (x, y) pairs (orientation irrelevant)
double polygon_inter_area(const double* a, int na, const double* b, int nb);
// |shoelace| area of a ring
double polygon_area(const double* a, int n);

// One find_matches call per map plane of one frame. w2c: row-major 4x4 world-to-camera. selected[m] = index of the matched
// detected plane inside the frame's detection list or -1; inter[m] = its intersection area (0 if none).
// sequential: the caller's loop too (Feature_Map::get_matches, feature_map.hpp:652-669) - a detection taken by a map plane is
// marked matched for the map planes after it; matched_out (n_det entries, may be null) = the mask as the loop leaves it.
void plane_match_frame(const double* w2c, const rs_polygon_plane* det, int n_det, const double* det_xy,
                       const rs_polygon_plane* map, int n_map, const double* map_xy, const unsigned char* det_matched,
                       int advanced_search, int sequential, int* selected, double* inter, unsigned char* matched_out);

}  // namespace oracle
