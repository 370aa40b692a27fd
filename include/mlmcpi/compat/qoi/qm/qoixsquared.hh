// qoixsquared.hh of eikehmueller/mlmcpathintegral (src/qoi/qm/qoixsquared.hh): the classes of this header live in
// mlmcpi_compat.hh, which puts the device-library implementations into the global namespace under
// the reference's names so that src/driver_qm.cc and src/driver_qft.cc compile unchanged.
#include "mlmcpi_compat.hh"
