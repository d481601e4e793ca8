// oracle/shim: boost::bind -> std::bind (used at Estimation/CellsDataContainer.cpp:265).
#pragma once
#include <functional>
namespace boost { using std::bind; }
using namespace std::placeholders;
