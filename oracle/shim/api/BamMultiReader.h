// oracle/shim: BamTools is absent; Estimation/CellsDataContainer.cpp:8 includes this header but uses nothing from it.
#pragma once
