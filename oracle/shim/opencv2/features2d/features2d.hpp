// ORACLE BUILD SHIM (test infrastructure): cv::KeyPoint lives in the core shim.
#include <opencv2/core/core.hpp>
