/* TEST INFRASTRUCTURE ONLY (oracle/): GSL status codes used by
 * action/qft/quenchedschwingerrenormalisation.cc:46-58. */
#ifndef MLMCPI_ORACLE_SHIM_GSL_ERRNO_H
#define MLMCPI_ORACLE_SHIM_GSL_ERRNO_H
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2 };
#endif
