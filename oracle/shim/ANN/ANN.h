/* TEST INFRASTRUCTURE ONLY (oracle shim).  Declaration-only stand-in so that the reference's icp_def.h, which
 * includes AnnPairAssignment.h, can be compiled; AnnPairAssignment itself is never instantiated by the node
 * (src/ThreadLocalize.cpp:211 uses FlannPairAssignment) and its .cpp is not part of the oracle build. */
#ifndef ORACLE_ANN_SHIM_H
#define ORACLE_ANN_SHIM_H
class ANNkd_tree;
typedef double ANNcoord;
typedef ANNcoord* ANNpoint;
typedef ANNpoint* ANNpointArray;
#endif
