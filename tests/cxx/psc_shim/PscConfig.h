// TEST INFRASTRUCTURE: stands in for the cmake-generated src/include/PscConfig.h (compile check only)
#pragma once
