// TEST INFRASTRUCTURE — estimator/parameters.h includes OpenCV for readParameters(); the factor sources use none of it.
#pragma once
