// Forwarding header: same include path as the reference (src/obcore/base/Logger.h); the classes live in obvious_b200.h.
#pragma once
#include "../../obvious_b200.h"
// The node silences the reference's logger (src/slam.cpp:17); the macros are kept as no-ops.
#define DBG_DEBUG 1
#define DBG_WARN 2
#define DBG_ERROR 3
#ifndef LOGMSG
#define LOGMSG(prio, msg) do { } while(0)
#define LOGMSG_CONF(file, conf, prioFile, prioScreen) do { } while(0)
#endif
