// TEST INFRASTRUCTURE: stands in for the cmake-generated libmrc/include/mrc_config.h (compile check only)
#pragma once
