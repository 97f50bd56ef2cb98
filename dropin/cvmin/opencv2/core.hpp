// forwards to the minimal OpenCV stand-in (see cvmin.h)
#pragma once
#include "../cvmin.h"
