/* TEST INFRASTRUCTURE ONLY (oracle shim), see ANN.h */
#include "ANN.h"
