#include <boost/unordered_map.hpp>
