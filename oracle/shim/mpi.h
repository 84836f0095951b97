// Minimal single-process stand-in for <mpi.h>: libgcm only reads rank/size
// (reference: src/libgcm/engine/GlobalVariables.hpp:69-73) and calls MPI_Init/Finalize in mains.
#pragma once
namespace MPI {
struct Comm {
	int Get_rank() const { return 0; }
	int Get_size() const { return 1; }
};
static Comm COMM_WORLD;
}
inline int MPI_Init(int*, char***) { return 0; }
inline int MPI_Finalize() { return 0; }
