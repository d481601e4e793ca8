/* TEST INFRASTRUCTURE ONLY (see stub_device.c): the one fact the untyped stubs need from the real header */
#include "../../include/dropest_b200.h"
const unsigned long stub_dge_config_size = sizeof(dge_config);
