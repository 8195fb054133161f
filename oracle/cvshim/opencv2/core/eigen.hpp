// oracle-build stand-in (see cvshim_core.hpp)
#include "../cvshim_core.hpp"
#ifndef OPB_CVSHIM_EIGEN
#define OPB_CVSHIM_EIGEN
#include <fstream>
namespace cv
{
// parsed by the reference's GlobalRegistration.cpp:16-28 (Eigen2OpenCV, a helper nothing calls)
template <class E> void eigen2cv(const E &, Mat &) {}
} // namespace cv
#endif
