// forwards to the cv::Mat stand-in (the catkin snapshot includes <opencv2/opencv.hpp>, ROS/lsd/include/myLSD.h:37)
#include "../opencv.hpp"
