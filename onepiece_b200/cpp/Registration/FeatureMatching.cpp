// Drop-in definitions of two steps of registration::RansacRegistration (reference src/Registration/GlobalRegistration.cpp,
// declared in GlobalRegistration.h:27-31) over the C-ABI:
//   FeatureMatching3D      :29-73   nearest target FPFH descriptor of every source descriptor, on the GPU
//   RejectMatchesRanSaPC   :75-108  distance-preservation test against randomly drawn matches; the library walks the caller's
//                                   std::default_random_engine exactly as libstdc++ would, so the engine leaves in the state
//                                   the reference would leave it in and the kept matches are the reference's
// A maintainer deletes these two bodies from GlobalRegistration.cpp (RansacRegistration and the GRANSAC estimator stay) and adds
// this file.
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <vector>

#include "Registration/GlobalRegistration.h"
#include "onepiece_b200.h"

namespace one_piece
{
namespace registration
{
namespace
{
opb_kdtree *Workspace()
{
    static opb_kdtree *ws = nullptr; // callers are single-threaded
    if (!ws && opb_kdtree_create(0, &ws) != OPB_OK)
    {
        std::cout << RED << "[ERROR]::[GlobalRegistration]::" << opb_last_error() << RESET << std::endl;
        std::exit(1); // no device: there is no CPU path
    }
    return ws;
}
void Flatten33(const FeatureSet &f, std::vector<float> &out)
{
    out.resize(f.size() * 33);
    for (size_t i = 0; i < f.size(); ++i)
        for (int k = 0; k < 33; ++k) out[33 * i + k] = (float)f[i](k);
}
void Flatten3(const geometry::Point3List &p, std::vector<float> &out)
{
    out.resize(p.size() * 3);
    for (size_t i = 0; i < p.size(); ++i)
        for (int k = 0; k < 3; ++k) out[3 * i + k] = (float)p[i](k);
}
} // namespace

void FeatureMatching3D(const FeatureSet &source_feature, const FeatureSet &target_feature, geometry::FMatchSet &matching_index)
{
    std::vector<float> sf, tf;
    Flatten33(source_feature, sf);
    Flatten33(target_feature, tf);
    std::vector<int32_t> pairs(2 * source_feature.size() + 2);
    size_t n = 0;
    matching_index.clear();
    if (opb_kdtree_feature_matching(Workspace(), sf.data(), source_feature.size(), tf.data(), target_feature.size(), pairs.data(), &n) != OPB_OK)
    {
        std::cout << RED << "[ERROR]::[FeatureMatching3D]::" << opb_last_error() << RESET << std::endl;
        return;
    }
    for (size_t i = 0; i < n; ++i) matching_index.push_back(std::make_pair(pairs[2 * i], pairs[2 * i + 1]));
}

void RejectMatchesRanSaPC(const geometry::Point3List &source_points, const geometry::Point3List &target_points,
                          std::default_random_engine &engine, geometry::FMatchSet &init_matches, int candidate_num, float difference)
{
    std::vector<float> s, t;
    Flatten3(source_points, s);
    Flatten3(target_points, t);
    std::vector<int32_t> pairs(2 * init_matches.size() + 2);
    for (size_t i = 0; i < init_matches.size(); ++i) { pairs[2 * i] = init_matches[i].first; pairs[2 * i + 1] = init_matches[i].second; }
    size_t n = init_matches.size();
    // the engine's whole state is its last output; the standard's stream operators are the portable way in and out
    std::ostringstream os;
    os << engine;
    uint32_t state = (uint32_t)std::stoul(os.str());
    if (opb_reject_matches(s.data(), source_points.size(), t.data(), target_points.size(), pairs.data(), &n, &state, candidate_num, difference) != OPB_OK)
    {
        std::cout << RED << "[ERROR]::[RejectMatchesRanSaPC]::" << opb_last_error() << RESET << std::endl;
        return;
    }
    std::istringstream is(std::to_string(state));
    is >> engine;
    init_matches.clear();
    for (size_t i = 0; i < n; ++i) init_matches.push_back(std::make_pair(pairs[2 * i], pairs[2 * i + 1]));
}
} // namespace registration
} // namespace one_piece
