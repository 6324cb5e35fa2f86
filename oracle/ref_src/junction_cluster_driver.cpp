// ORACLE build input (test infrastructure only).  Driver around the reference's OWN vendored nanoflann header
//   /root/reference/ros2_ws/src/junction_point_detector/include/junction_point_detector/vendor/nanoflann/nanoflann.hpp
// (included from where it lies, never copied): reads junction candidates "x y" per line from stdin, runs the
// clustering step of find_junctions_not_rotated (junction_detector.cpp:129-185: KD-tree with leaf size 7 built twice,
// radiusSearch(radius^2, SearchParameters(10.0, false)), clusters of >= 3 neighbours, greedy `visited` marking in
// candidate order, centre = mean) and prints the cluster centres with %.9g.  junction_detector.cpp itself cannot be
// compiled here (it needs the OpenCV C++ headers, which this image does not have — only the cv2 wheel), so the
// 50-line loop is restated in this driver; what the binary pins is nanoflann's tree build and its APPROXIMATE
// (eps = 10) radius search, which the clustering result depends on.
// usage: junction_cluster <eps_radius>  < points.txt
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nanoflann.hpp"

struct Pt { float x, y; };
typedef std::vector<Pt> Cloud;

// same contract as the reference's KDTreeVectorOfCVPoint2fAdaptor.h (num_t float, DIM 2, metric_L2, size_t index)
struct Adaptor {
  using self_t = Adaptor;
  using metric_t = nanoflann::metric_L2::traits<float, self_t>::distance_t;
  using index_t = nanoflann::KDTreeSingleIndexAdaptor<metric_t, self_t, 2, size_t>;
  index_t* index = nullptr;
  const Cloud& m_data;
  Adaptor(const Cloud& c, int leaf_max_size) : m_data(c) {
    index = new index_t(2, *this, nanoflann::KDTreeSingleIndexAdaptorParams(leaf_max_size, nanoflann::KDTreeSingleIndexAdaptorFlags::None, 1));
  }
  ~Adaptor() { delete index; }
  inline size_t kdtree_get_point_count() const { return m_data.size(); }
  inline float kdtree_get_pt(const size_t idx, const size_t dim) const { return dim == 0 ? m_data[idx].x : m_data[idx].y; }
  template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};

int main(int argc, char** argv) {
  const int eps = argc > 1 ? atoi(argv[1]) : 6;
  Cloud junctions;
  float x, y;
  while (scanf("%f %f", &x, &y) == 2) junctions.push_back(Pt{x, y});
  if (junctions.size() < 4) return 0;
  Adaptor index(junctions, 7);
  index.index->buildIndex();
  const float radius = eps;
  std::vector<bool> visited(junctions.size());
  nanoflann::SearchParameters params(10.0, false);
  for (size_t i = 0; i < junctions.size(); ++i) {
    if (visited[i]) continue;
    std::vector<nanoflann::ResultItem<size_t, float>> neighbors;
    index.index->radiusSearch(&junctions[i].x, radius * radius, neighbors, params);
    if (neighbors.size() >= 3) {
      float cx = 0, cy = 0;
      for (const auto& nb : neighbors) { cx += junctions[nb.first].x; cy += junctions[nb.first].y; }
      cx /= neighbors.size();
      cy /= neighbors.size();
      printf("%.9g %.9g\n", cx, cy);
      for (const auto& nb : neighbors) visited[nb.first] = true;
    }
  }
  return 0;
}
