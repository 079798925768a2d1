// TEST INFRASTRUCTURE ONLY (oracle build). Single-rank stand-in for the Boost.MPI
// surface the reference touches (src/utility.h:21-27, src/newforces.cpp:268-271,
// src/interface.cpp:1092-1094). With size()==1 the reference never reaches the
// collectives, so these bodies only have to compile.
#ifndef NPSL_SHIM_BOOST_MPI_HPP
#define NPSL_SHIM_BOOST_MPI_HPP
#include <vector>
namespace boost {
namespace serialization { class access; }
namespace mpi {
struct environment { environment() {} };
struct communicator {
    int rank() const { return 0; }
    int size() const { return 1; }
};
template <class T>
inline void all_gather(const communicator &, const T *in, int n, std::vector<T> &out) {
    out.assign(in, in + n);
}
template <class T, class Op>
inline void all_reduce(const communicator &, const T &in, T &out, Op) { out = in; }
}  // namespace mpi
}  // namespace boost
#endif
