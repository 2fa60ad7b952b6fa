// Forwarding header: same include path as the reference (src/obvision/reconstruct/grid/RayCastAxisAligned2D.h); the classes live in obvious_b200.h.
#pragma once
#include "../../../obvious_b200.h"
