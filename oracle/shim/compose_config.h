/* TEST INFRASTRUCTURE ONLY (oracle/). Stand-in for the file CMake would
 * generate from /root/reference/compose_config.h.in. All three switches
 * (COMPOSE_DEBUG_MPI, COMPOSE_MIMIC_GPU, COMPOSE_QLT_TIME) are left off:
 * the baseline is the reference's plain host path. */
#ifndef COMPOSE_CONFIG_H
#define COMPOSE_CONFIG_H
#endif
