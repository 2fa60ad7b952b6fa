// Forwarding header: same include path as the reference (src/obvision/registration/ransacMatching/PDFMatching.h); the classes live in obvious_b200.h.
#pragma once
#include "../../../obvious_b200.h"
