// Stand-in for <boost/unordered_map.hpp> (Boost is not installed here).  TEST INFRASTRUCTURE, see Eigen/Dense.
// An ordered map keyed by operator< serves: the reference never depends on the iteration order (it only appends the
// per-key vectors of one map to the same keys of another, HFTest.cpp:648-652).
#ifndef HF6D_SHIM_BOOST_UNORDERED_MAP
#define HF6D_SHIM_BOOST_UNORDERED_MAP
#include <cstddef>
#include <functional>
#include <map>
#include <set>

namespace boost {
template <typename T>
struct hash {
    std::size_t operator()(const T& v) const { return (std::size_t)v; }
};
template <typename It>
std::size_t hash_range(It first, It last) {
    std::size_t seed = 0;
    for (; first != last; ++first) seed ^= (std::size_t)(*first) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
}
template <typename K, typename V, typename H = boost::hash<K>, typename E = std::equal_to<K> >
class unordered_map : public std::map<K, V> {};
template <typename K, typename H = boost::hash<K>, typename E = std::equal_to<K> >
class unordered_set : public std::set<K> {};
}  // namespace boost
#endif
