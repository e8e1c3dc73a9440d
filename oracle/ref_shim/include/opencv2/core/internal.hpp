// oracle/ref_shim: everything lives in opencv2/core/core.hpp
#include "opencv2/core/core.hpp"
