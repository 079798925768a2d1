#include <boost/mpi.hpp>
