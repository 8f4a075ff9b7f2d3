// csrc/nccl_dl.hpp -- NCCL entered through dlopen, so that libmf.so has no link-time dependency on it:
// a single-GPU host (the PHP extension) never needs NCCL installed, and a multi-GPU host process that
// already carries an NCCL (PyTorch's bundled one) shares that copy instead of loading a second one.
// Only the types come from <nccl.h>.
#ifndef MFB200_NCCL_DL_HPP
#define MFB200_NCCL_DL_HPP

#include <cuda_runtime.h>
#include <nccl.h>

namespace mfb200 {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
};

// nullptr (with set_error) when no NCCL library can be loaded.
const NcclApi *nccl_api();

}  // namespace mfb200
#endif
