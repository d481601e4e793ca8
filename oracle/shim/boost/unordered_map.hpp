// oracle/shim: boost::unordered_map -> std::unordered_map (typedef only, Tools/ReadParameters.h:52).
#pragma once
#include <unordered_map>
namespace boost { template <class K, class V> using unordered_map = std::unordered_map<K, V>; }
