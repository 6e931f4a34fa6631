#pragma once
#include "common/eigen_types.h"
