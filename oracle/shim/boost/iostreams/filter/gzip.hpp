// oracle/shim -- see ../filtering_stream.hpp
#pragma once
#include "../filtering_stream.hpp"
