#pragma once
#include <boost/algorithm/string.hpp>
