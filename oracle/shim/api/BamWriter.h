// oracle/shim -- TEST INFRASTRUCTURE ONLY.  BamTools is absent; the reference's BamProcessor headers only name these types.
#pragma once
#include "BamAlignment.h"
namespace BamTools { class BamWriter {}; }
