/* symgpu.h — thin C-ABI between the SymuVia-compatible host library and the sm_100a kernels.
 *
 * This is the seam SURVEY.md §8(b) places inside Reseau::SimuTrafic
 * (symuvia/src/reseau.cpp:2250-2258): everything from InitVehicules (:3901) to
 * SupprimeVehicules (:3984) runs on the device behind `symgpu_step`; vehicle creation at the
 * origins (usage/SymuViaTripNode.cpp:212-325) stays on the host and feeds rows in, the
 * trajectory rows the XML writers consume (Vehicule::SortieTrafic, vehicule.cpp:1993) come out.
 * Plain pointers and sizes only; no CPU fallback exists: every entry point fails with
 * SYMGPU_E_NODEVICE when no CUDA device is usable.
 */
#ifndef SYMGPU_H
#define SYMGPU_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct symgpu_ctx symgpu_ctx;

enum { SYMGPU_OK = 0, SYMGPU_E_NODEVICE = -1, SYMGPU_E_LOAD = -2, SYMGPU_E_CUDA = -3, SYMGPU_E_CAPACITY = -4,
       SYMGPU_E_UNSUPPORTED = -5, SYMGPU_E_ARG = -6 };

/* counters readable through symgpu_counter() */
enum { SYMGPU_CNT_TIES = 0,      /* equal-position leader candidates met (reseau.cpp:14041-14128 tie-break not replayed) */
       SYMGPU_CNT_LOOPS = 1,     /* leader-recursion loops broken (vehicule.cpp:1201-1216) */
       SYMGPU_CNT_CREATED = 2,   /* vehicles generated at origins */
       SYMGPU_CNT_PASSES = 3,    /* relaxation sweeps that changed at least one speed (flow.cuh) */
       SYMGPU_CNT_LAUNCHES = 4,  /* kernels launched since creation */
       SYMGPU_CNT_SLOTS = 5,     /* vehicle slots in use (alive + holes) */
       SYMGPU_CNT_CTX_OVERFLOW = 6, /* context lane list longer than the device bound (must stay 0) */
       SYMGPU_CNT_LANE_CHANGES = 7, /* successful lane changes (Vehicule::ChangementVoie) */
       SYMGPU_CNT_LC_EVENTS = 8,    /* lane-change events that drew or changed something (walked one by one) */
       SYMGPU_CNT_LC_REEVAL = 9,    /* acceptance-test records re-evaluated after a commit */
       SYMGPU_CNT_MRG_INS = 10,     /* merge points: insertions (PointDeConvergence::CalculInsertion accepted) */
       SYMGPU_CNT_MRG_REFUS = 11,   /* merge points: insertions refused (vehicle held at the end of its branch) */
       SYMGPU_CNT_MRG_SERIAL = 12,  /* merge points processed one by one because they shared a vehicle with another point */
       SYMGPU_CNT_MRG_VALUES = 13,  /* merge points whose decision consumed values of the random stream (congested branch) */
       SYMGPU_CNT_MRG_DRAWS = 14    /* draws of the random stream made by the merge phase */ };

/* Replaces Reseau::LoadReseauXML + InitSimuTrafic for an already flattened network (include/symflat.h). */
int symgpu_create(const char* symflat_path, int device, symgpu_ctx** out);
void symgpu_destroy(symgpu_ctx* c);
const char* symgpu_last_error(void);

/* One Reseau::SimuTrafic time step (reseau.cpp:2028-2428): returns vehicles on the network after the step. */
int symgpu_step(symgpu_ctx* c);
/* n_steps steps back to back; no per-step host copies unless the network has origins. */
int symgpu_run(symgpu_ctx* c, int n_steps);
/* One step + trajectory rows copied into caller (host) buffers of capacity `cap` rows: the data
 * Vehicule::SortieTrafic (vehicule.cpp:1993-2176) hands to DocTrafic::AddTrajectoire.  Returns rows written. */
int symgpu_step_traj(symgpu_ctx* c, int cap, int* id, int* lane, double* pos, double* vit, double* acc);

/* Asynchronous form (north_star: "trajectory output is copied back asynchronously on a side stream"): bind caller buffers once
 * (they are page-locked in place), then rotate them; the rows of step k are packed and copied on a side stream beside step k+1.
 * `which` = 0..3: two buffers let the consumer lag one step, three in rotation two. */
int symgpu_traj_bind(symgpu_ctx* c, int which, int cap, int* id, int* lane, double* pos, double* vit, double* acc);
int symgpu_step_traj_async(symgpu_ctx* c, int which);   /* returns rows, valid after symgpu_traj_wait(which) */
int symgpu_traj_wait(symgpu_ctx* c, int which);
double symgpu_traj_copy_ms(symgpu_ctx* c, int which);    /* device->host copy time of the rows last waited for (side stream, CUDA events) */
int symgpu_traj_copy_async(symgpu_ctx* c, int which);   /* the copy alone (current state), e.g. after symgpu_mp_import */

/* Partitioned runs: one process per GPU, network tables replicated, vehicles owned by the rank that owns their lane.
 * The caller moves the device buffers between ranks (NCCL send/recv through torch.distributed in the Python mirror).
 * A follower whose leader is the tail of a lane owned by another rank uses the reference's estimated leader speed
 * (NewellCarFollowing.cpp:217-230) instead of the leader's new speed: see DESIGN.md "multi-GPU". */
int symgpu_mp_configure(symgpu_ctx* c, int rank, int world, const int* node_part);
int symgpu_mp_sizes(symgpu_ctx* c, int* export_cnt, int* halo_cnt);                 /* tail records per peer (4 doubles each) */
int symgpu_mp_step_front(symgpu_ctx* c, double* send_tails);
int symgpu_mp_step_back(symgpu_ctx* c, const double* recv_tails, double* send_rows, int cap_per_peer, int* send_counts);
int symgpu_mp_import(symgpu_ctx* c, const double* recv_rows, int n_rows);          /* rows of 16 doubles */
int symgpu_mp_import_regions(symgpu_ctx* c, const double* recv_regions, int cap_per_peer);   /* the receive buffer of the fixed-size all_to_all as it is (world regions of cap_per_peer rows) */
/* Tail exchange over peer memory (NVLink, CUDA IPC) instead of a collective: _handle returns this rank's 64-byte IPC handle (and its
   halo size); _connect takes the world's handles, and per peer q the offset (in records) of this rank's lanes inside q's halo and q's
   halo size.  Afterwards symgpu_mp_step_front stores the tails into the peers and symgpu_mp_step_back waits for theirs on the device;
   send_tails / recv_tails are ignored. */
int symgpu_mp_p2p_handle(symgpu_ctx* c, void* handle64);
int symgpu_mp_p2p_connect(symgpu_ctx* c, const void* handles, const int* peer_halo_off, const int* peer_n_halo);
int symgpu_set_stream(symgpu_ctx* c, void* cuda_stream);   /* run on the caller's stream (cudaStream_t), e.g. the one its collectives are ordered on */
/* Origins in a partitioned run: every rank replays the creation draws of all origins (same seed, same ids) and keeps the vehicles of
   the entry lanes it owns; the device's own draws of the step are summed over the ranks by the caller and committed on every rank. */
int symgpu_mp_has_origins(symgpu_ctx* c);
long symgpu_mp_local_draws(symgpu_ctx* c);
int symgpu_mp_commit_draws(symgpu_ctx* c, long total_over_ranks);

/* Per-kernel device timing (CUDA events on the step stream) for roofline accounting.  enable = 1: every kernel group of a step is
   bracketed (breakdown); 2: only the relaxation sweeps (the dominant kernel, timed inside a measured region); 0: off. */
int symgpu_profile(symgpu_ctx* c, int enable);
int symgpu_kernel_count(void);
const char* symgpu_kernel_name(int k);
double symgpu_kernel_ms(symgpu_ctx* c, int k);
long symgpu_kernel_launches(symgpu_ctx* c, int k);
long long symgpu_vehicle_steps(symgpu_ctx* c);   /* sum over steps of vehicles on the network after the step */

int symgpu_num_vehicles(symgpu_ctx* c);
double symgpu_time(symgpu_ctx* c);
unsigned symgpu_rng_count(symgpu_ctx* c);       /* RandManager::m_Count (RandManager.cpp:26-30) */
long symgpu_counter(symgpu_ctx* c, int which);
double symgpu_last_step_ms(symgpu_ctx* c);      /* device time of the last step (CUDA events on the step stream) */

/* State of m_LstVehicles (list order) after the last step; any pointer may be NULL. Returns rows. */
int symgpu_get_state(symgpu_ctx* c, int cap, int* id, int* lane, int* route_pos, int* flags, double* pos, double* vit,
                     double* acc, double* deltaN, int* leader_id);
/* Tuyau::m_mapNbVehEntre / m_mapNbVehSorti (tuyau.h:531-532), arrays of n_links * n_classes. */
int symgpu_get_link_counts(symgpu_ctx* c, int* entre, int* sorti);
/* Vehicles removed by Reseau::SupprimeVehicules (reseau.cpp:3984) during the last step. */
int symgpu_get_exited(symgpu_ctx* c, int cap, int* id, double* instant);
int symgpu_num_links(symgpu_ctx* c);
int symgpu_num_classes(symgpu_ctx* c);
int symgpu_waiting(symgpu_ctx* c);              /* vehicles queued at origins (m_mapVehEnAttente) */
/* Merge points (PointDeConvergence, convergent.h) in Reseau::Liste_convergents order: m_dbInstLastPassage and m_nRandNumbers.
   Returns the number of points; fills at most `cap`. */
int symgpu_get_merge_state(symgpu_ctx* c, int cap, double* inst_last_passage, int* rand_numbers);

#ifdef __cplusplus
}
#endif
#endif
