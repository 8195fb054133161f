// TEST BUILD ONLY: link stand-ins for the sparse-tracking members that one_piece::odometry::Odometry's inline
// constructors reference (MILD's SparseMatcher lives in 3rdparty/MILD and needs real OpenCV).  Never called.
#include "Odometry/Odometry.h"
namespace MILD
{
SparseMatcher::SparseMatcher(int, int, int, float) {}
SparseMatcher::~SparseMatcher() {}
} // namespace MILD
