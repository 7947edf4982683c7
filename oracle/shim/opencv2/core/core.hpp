/* opencv2/core/core.hpp stand-in: /root/reference/include/JointWMF.h:38 includes this path. See ../opencv.hpp. */
#include "../opencv.hpp"
