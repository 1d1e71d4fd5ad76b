// One-shot all-reduce of small replicated vectors over NVLink peer memory, fused with the second
// stage of the local reduction that produces them.
//
// Every rank owns an exchange buffer (cudaMalloc + cudaIpc handle, mapped by all peers):
//   flags[r]      sequence number of the latest exchange rank r has published here
//   data[2][cap]  two slots, used alternately
// One kernel per exchange and rank:
//   (0) every CTA reads the sequence number of this exchange from the device counter (+1);
//   (1) fold this rank's per-CTA partials (or copy a ready vector) straight into the local slot,
//       fence at system scope, and let the last CTA announce `seq` in every peer's flag word with
//       a remote store;
//   (2) wait until every rank's flag here reached `seq`, then add the W slots (own + peers',
//       remote loads issued in parallel) in rank order -- every rank adds in the same order, so
//       the result is bit-identical everywhere -- and, for the stop test, update the control block.
// Two slots are enough: finishing exchange k+1 needs every peer's flag k+1, which a peer only
// raises after it finished reading slot k.  Exchange kernels are never skipped (not even in the
// empty trips enqueued past convergence), so sequence numbers and slots stay aligned on all ranks.
// Replaces reduce_cols + ncclAllReduce (+ stop) on the per-trip critical path.
#pragma once

#include "passes.cuh"
#include "small.cuh"

namespace tpls {

constexpr int kXchgMaxRanks = 16;
constexpr size_t kXchgHeaderBytes = 1024;
// header of an exchange buffer: flags[kXchgMaxRanks] (u64) | +512 done counter (u32) | +520 error word (i32) |
// +528 sequence number of the last exchange this rank completed (u64; kept on the DEVICE so that an exchange can
// sit in the body of a CUDA-graph loop) | +536 diagnostics (u64 x 3, reset by every fit): nanoseconds CTA 0 spent
// waiting for the peers' announcements, nanoseconds from its start to the end of that wait, exchanges counted

struct XchgArgs {
    // input: either `n_sets` partial blocks (folded into the slot at their `off`), or (n_sets == 0) the ready vector `in`
    FoldSet sets[kMaxFoldSets];
    int n_sets;
    const double* in;
    double* out;             // summed vector [count] (may alias in)
    int count;
    int cap;                 // doubles per slot
    int rank, world;
    unsigned long long* seq_ctr;               // local: exchanges completed so far (identical on all ranks)
    unsigned long long* flags[kXchgMaxRanks];  // flags[r] = rank r's flag array (flags[rank] is local)
    double* data[kXchgMaxRanks];               // data[r] = rank r's slots
    unsigned int* done_ctr;                    // local: CTAs that finished phase 1
    int* err;                                  // local: set to 1 on a wait timeout
    unsigned long long* diag;                  // local: {wait ns, start-to-synchronised ns, exchanges}
    // optional stop test on out[0] = ||u_old - u_new||^2 (tpls.py:103), see ctrl_decide
    Ctrl* ctrl;
    LoopEnd loop_end;
    int do_stop;
    // optional (count <= 32, one CTA): `out` is the raw q = Y't -- normalise it and run the stop test through
    // dq^T (Y'Y) dq in this kernel (small.cuh normalize_q_stop_body); uses ctrl / loop_end above
    int do_qstop;
    int q_m, q_pitch;
    double* qcol;
    double* qvec;
    const double* gram;
    double* q_prev;
};

cudaError_t launch_xchg(const XchgArgs& a, cudaStream_t s);

}  // namespace tpls
