/* b200comm.h — the one collective of the batched loop-closure sweep, from C (north_star: "a single
 * NCCL all-gather of the resulting 4x4 poses"): every rank registers its share of the (scan, submap) pairs
 * (b200reg_ndt_sweep, include/b200reg.h) and ONE ncclAllGather makes all result rows visible to every rank — and to the host
 * that feeds the pose graph (graph_based_slam_component.cpp:236-258 builds the loop edges from them).
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2 — the copy a host application already loaded, e.g. PyTorch's, is
 * reused), so libb200reg.so has no link-time dependency on it and single-GPU users never load it. The 128-byte unique id
 * travels out of band (the launcher's rendezvous: MPI, a TCP store, ROS parameters ...), exactly like ncclGetUniqueId /
 * ncclCommInitRank expect. One communicator = one rank = one GPU.                                                          */
#ifndef B200COMM_H_
#define B200COMM_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200comm* b200comm_t;
#define B200COMM_UNIQUE_ID_BYTES 128

int b200comm_unique_id(unsigned char out[B200COMM_UNIQUE_ID_BYTES]);            /* rank 0: ncclGetUniqueId          */
int b200comm_create(const unsigned char id[B200COMM_UNIQUE_ID_BYTES], int rank, int world, int device, b200comm_t* out);
int b200comm_destroy(b200comm_t c);
/* All-gather of fixed-size float rows: every rank contributes rows_per_rank rows of row_floats floats (host memory);
 * rows_all (host, world * rows_per_rank * row_floats floats) receives them in rank order. One pinned H2D copy, ONE
 * ncclAllGather on the communicator's own stream, one D2H copy; returns when rows_all is complete.                  */
int b200comm_all_gather_rows(b200comm_t c, const float* rows_local, int rows_per_rank, int row_floats, float* rows_all);
int b200comm_rank(b200comm_t c, int* rank, int* world);

/* Pose board — the exchange fused INTO the registration kernel. For the sharded batch of one map (every rank registers
 * its own scans against the same submap: b200reg_ndt_align_batch[_device]) the 4x4 poses do not need a collective call
 * at all: each rank owns a small board in HBM, all ranks map all boards (CUDA IPC over NVLink / NVSwitch peer memory),
 * and the solver kernel's controller CTAs store every finished pose straight into all boards, tagged word by word
 * with the launch number. The transfer overlaps the remaining registrations of the launch; when the call returns,
 * b200reg_ndt_gathered_poses (b200reg.h) has every rank's poses. Creation is collective over the communicator (one
 * all-gather of the 64-byte IPC handles); max_rows = the largest batch a rank will register in one call. While a
 * board is attached to a handle (b200reg_ndt_attach_pose_board), that handle's batch calls are collective: every rank
 * of the board must make the same sequence of them.                                                                 */
typedef struct b200comm_board* b200comm_board_t;
int b200comm_board_create(b200comm_t c, int max_rows, b200comm_board_t* out);
int b200comm_board_destroy(b200comm_board_t b);
int b200comm_board_info(b200comm_board_t b, int* rank, int* world, int* max_rows);
const char* b200comm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* B200COMM_H_ */
