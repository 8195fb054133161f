// TEST BUILD ONLY: headless stand-in for the reference's src/Visualization/Visualizer.h (Pangolin / OpenGL window), so that the
// reference's example mains compile and run UNMODIFIED on a machine without a display.  Every call is a no-op; nothing here is
// part of the product.  Method names follow the reference class (Visualizer.h:19-120).
#ifndef OPB_HEADLESS_VISUALIZER_H
#define OPB_HEADLESS_VISUALIZER_H
#include <string>

#include "Geometry/Geometry.h"
#include "Geometry/PointCloud.h"
#include "Geometry/TriangleMesh.h"
namespace one_piece
{
namespace visualization
{
class Visualizer
{
  public:
    void AddPointCloud(const geometry::PointCloud &) {}
    void AddTriangleMesh(const geometry::TriangleMesh &) {}
    void AddCameraSet(const geometry::SE3List &, const geometry::Point3List &) {}
    void Show() {}
    void ShowOnce() {}
    void Reset() {}
    void Initialize(const std::string & = "OnePiece") {}
    void DrawPhongRendering() {}
    void SetDrawColor(bool) {}
};
} // namespace visualization
} // namespace one_piece
#endif
