// forwards to the cv::Mat stand-in (the catkin snapshot includes <opencv2/core.hpp>, ROS/lsd/include/FeatureAssociation.h:6)
#include "../opencv.hpp"
