/*
 * dvp_mvs.h — flat C ABI of the B200-native PatchMatch-MVS engine (libdvp_mvs.so).
 *
 * This is the drop-in boundary for ONE path of ZhenlongYuan/DVP-MVS: APD::RunPatchMatch()
 * and the 16 CUDA kernels it launches (reference APD.cu:4406-4532).  The reference has no
 * FFI/plugin layer; its boundary is the C++ class `APD` (reference APD.h:94-199) as driven by
 * ProcessProblem (reference main.cpp:267-419).  Every entry point below names the reference
 * interface it replaces.  `include/dvp_apd_adapter.hpp` is the header-only `APD` class a
 * maintainer compiles main.cpp against.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch/OpenCV types cross this boundary.
 *  - all image-shaped arrays are row-major, index y*W + x, W/H = reference-view size.
 *  - every function returns 0 on success or a negative dvp_status; CUDA errors are returned as
 *    DVP_ERR_CUDA with the cudaError_t retrievable through dvp_last_cuda_error().  Nothing in the
 *    library calls exit() (the reference's CUDA_SAFE_CALL does, APD.cpp:943-951; the adapter
 *    restores that behaviour for main.cpp).
 *  - a context is bound to one device and owns one stream; contexts on different devices may be
 *    driven from different host threads concurrently (the 8-GPU view farm relies on this).
 */
#ifndef DVP_MVS_H
#define DVP_MVS_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVP_MAX_IMAGES 32      /* reference main.h:39  MAX_IMAGES      */
#define DVP_NEIGHBOUR_NUM 12   /* reference main.h:40  NEIGHBOUR_NUM   */
#define DVP_NUM_IMAGES 4       /* reference main.h:41  NUM_IMAGES      */
#define DVP_EDGE_NEIGH_NUM 8   /* reference main.h:43  EDGE_NEIGH_NUM  */
#define DVP_LAB_BOUNDARY_NUM 8 /* reference main.h:44  LAB_BOUNDARY_NUM*/

typedef enum dvp_status {
	DVP_OK = 0,
	DVP_ERR_ARG = -1,      /* null pointer / bad size / bad id                         */
	DVP_ERR_CUDA = -2,     /* a CUDA runtime call failed; see dvp_last_cuda_error()    */
	DVP_ERR_STATE = -3,    /* call order violated (e.g. run before upload)             */
	DVP_ERR_UNSUPPORTED = -4
} dvp_status;

/* reference main.h:74-78 RunState */
enum { DVP_FIRST_INIT = 0, DVP_REFINE_INIT = 1, DVP_REFINE_ITER = 2 };
/* reference main.h:80-84 PixelState */
enum { DVP_WEAK = 0, DVP_STRONG = 1, DVP_UNKNOWN = 2 };

/* reference main.h:58-67 `struct Camera` — identical byte layout (112 B). */
typedef struct dvp_camera {
	float K[9];
	float R[9];
	float t[3];
	float c[3];
	int32_t height;
	int32_t width;
	float depth_min;
	float depth_max;
} dvp_camera;

/* reference main.h:86-112 `struct PatchMatchParams`, one field per reference field, same order.
 * bools are widened to int32 so the struct has one layout in C, C++ and ctypes. */
typedef struct dvp_params {
	int32_t max_iterations;    /* 3     */
	int32_t num_images;        /* 1 + S */
	float sigma_spatial;       /* 5.0   */
	float sigma_color;         /* 3.0   */
	int32_t top_k;             /* 4     */
	float depth_min;
	float depth_max;
	int32_t geom_consistency;
	int32_t strong_radius;     /* 5 */
	int32_t strong_increment;  /* 2 */
	int32_t weak_radius;       /* 5 */
	int32_t weak_increment;    /* 5 */
	int32_t use_APD;
	int32_t use_edge;          /* must be 1: with 0 the reference reads an uninitialised position array (APD.cu:2036, 2559) -> DVP_ERR_UNSUPPORTED */
	int32_t use_limit;
	int32_t use_label;
	int32_t use_detail;
	int32_t use_radius;
	int32_t weak_peak_radius;  /* 2 */
	int32_t rotate_time;       /* 4 */
	float ransac_threshold;    /* 0.005 */
	float geom_factor;         /* 0.2   */
	int32_t state;             /* DVP_FIRST_INIT / DVP_REFINE_INIT / DVP_REFINE_ITER */
} dvp_params;

/* Host-side inputs of one RunPatchMatch call == what InuputInitialization + SupportInitialization
 * leave in the APD object (reference APD.cpp:1045-1495, 1615-1668) and CudaSpaceInitialization
 * uploads (APD.cpp:1497-1613).  Optional pointers may be NULL. */
typedef struct dvp_inputs {
	const float* images;          /* [(1+S)][H][W] f32 grey levels; image 0 = reference view           */
	const float* depths;          /* [(1+S)][H][W] f32 or NULL; required iff params.geom_consistency   */
	const dvp_camera* cameras;    /* [1+S]                                                             */
	const float* planes;          /* [H][W][4] (world nx,ny,nz, depth) — plane_hypotheses_host         */
	const uint32_t* selected_views; /* [H][W] bit i = source i+1, or NULL = all zero (FIRST_INIT)      */
	const uint8_t* weak_info;     /* [H][W] DVP_WEAK/STRONG/UNKNOWN or NULL = all STRONG (use_APD=0)   */
	const uint8_t* edge;          /* [H][W] 0/255, or NULL = no edges                                  */
	const int32_t* label;         /* [H][W] region label (0 boundary, -1 small, >0 id) or NULL = zeros */
	const int32_t* radius;        /* [H][W] NCC patch radius or NULL = strong_radius everywhere        */
	uint64_t seed;                /* cuRAND XORWOW seed (the reference uses clock64(), APD.cu:1270)    */
} dvp_inputs;

/* Device buffers addressable through dvp_get_buffer / dvp_set_buffer (parity stepping).
 * Element layout is the reference's except DVP_BUF_RAND, exchanged as 6 x u32
 * per pixel {d, v0, v1, v2, v3, v4} (the live part of curandStateXORWOW). */
typedef enum dvp_buffer {
	DVP_BUF_PLANES = 0,        /* float4  [N]        plane_hypotheses_cuda          */
	DVP_BUF_COSTS = 1,         /* float   [N]        costs_cuda (never leaves the device in the reference) */
	DVP_BUF_SELECTED = 2,      /* uint32  [N]        selected_views_cuda            */
	DVP_BUF_WEAK = 3,          /* uint8   [N]        weak_info_cuda                 */
	DVP_BUF_RADIUS = 4,        /* int32   [N]        radius_cuda                    */
	DVP_BUF_VIEW_WEIGHT = 5,   /* uint8   [N][32]    view_weight_cuda               */
	DVP_BUF_RAND = 6,          /* uint32  [N][6]     rand_states_cuda (canonical)   */
	DVP_BUF_FIT_PLANES = 7,    /* float4  [N]        fit_plane_hypotheses_cuda      */
	DVP_BUF_EDGE_NEIGH = 8,    /* short2  [N][8]     edge_neigh_cuda                */
	DVP_BUF_CANDIDATE = 9,     /* short2  [N][4][8]  candidate_cuda                 */
	DVP_BUF_NEAREST_STRONG = 10, /* short2 [N]       weak_nearest_strong            */
	DVP_BUF_WEAK_RELIABLE = 11,  /* uint8  [N]       weak_reliable_cuda             */
	DVP_BUF_NEIGHBOURS_MAP = 12, /* int32  [N]       neighbours_map_cuda            */
	DVP_BUF_NEIGHBOURS = 13,   /* short2  [weak_count][12]  neighbours_cuda         */
	DVP_BUF_LABEL_BOUNDARY = 14, /* short2 [weak_count][8]  label_boundary_cuda     */
	DVP_BUF_COMPLEX = 15,      /* float   [weak_count]      complex_cuda            */
	DVP_BUF_COUNT = 16
} dvp_buffer;

/* One id per kernel launched by the reference RunPatchMatch, in launch order (APD.cu:4430-4505). */
typedef enum dvp_stage {
	DVP_K1_INIT_RANDOM_STATES = 0,   /* APD.cu:1258 */
	DVP_K2_GEN_EDGE_INFORM = 1,      /* APD.cu:3731 */
	DVP_K3_FIND_NEAREST_STRONG = 2,  /* APD.cu:4159 */
	DVP_K4_GEN_NEIGHBOURS = 3,       /* APD.cu:3330 */
	DVP_K5_NEIGHBOUR_UPDATE = 4,     /* APD.cu:3713 */
	DVP_K6_RANDOM_INITIALIZATION = 5,/* APD.cu:1273 */
	DVP_K7_BLACK_STRONG = 6,         /* APD.cu:3127 (iter) */
	DVP_K8_RED_STRONG = 7,           /* APD.cu:3147 (iter) */
	DVP_K9_RANSAC_FIT_PLANE = 8,     /* APD.cu:4195 */
	DVP_K10_BLACK_WEAK = 9,          /* APD.cu:3091 (iter) */
	DVP_K11_RED_WEAK = 10,           /* APD.cu:3109 (iter) */
	DVP_K12_DEPTH_NORMAL = 11,       /* APD.cu:3167 */
	DVP_K13_BLACK_FILTER = 12,       /* APD.cu:3296 */
	DVP_K14_RED_FILTER = 13,         /* APD.cu:3313 */
	DVP_K15_DEPTH_TO_WEAK = 14,      /* APD.cu:3892 */
	DVP_K16_LOCAL_REFINE = 15,       /* APD.cu:4053 */
	DVP_STAGE_COUNT = 16,
	/* product only: K15 and K16 in one launch, as dvp_run issues them (accepted by dvp_run_stage for parity tests) */
	DVP_K15_K16_FUSED = 16
} dvp_stage;

typedef struct dvp_ctx dvp_ctx;

/* Library identity: "dvp_mvs_b200 <version> sm_100a" (the oracle harness answers "reference"). */
const char* dvp_version(void);

/* Fill *p with the reference defaults (main.h:86-112). */
void dvp_default_params(dvp_params* p);

/* Replaces: APD::APD(problem) + the cudaMalloc half of CudaSpaceInitialization (APD.cpp:984, 1497-1613).
 * Allocates every device buffer for a W x H reference view with S source views on `device`.
 * Returns NULL on failure. */
dvp_ctx* dvp_create(int device, int width, int height, int num_src, const dvp_params* params);

/* Replaces: ~APD (APD.cpp:989-1043). */
void dvp_destroy(dvp_ctx* ctx);

/* Replaces: the H2D half of CudaSpaceInitialization + SetDataPassHelperInCuda (APD.cpp:1497-1613,
 * 1670-1704).  `params` may change between uploads on the same context (multi-pass reuse).
 * Host pointers may be pageable or pinned; copies are issued on the context stream. */
int dvp_upload(dvp_ctx* ctx, const dvp_inputs* in, const dvp_params* params);

/* Replaces: APD::RunPatchMatch kernel sequence (APD.cu:4430-4505), all stages, one stream, no host
 * synchronisation between stages.  Asynchronous unless `sync` != 0. */
int dvp_run(dvp_ctx* ctx, int sync);

/* One stage of the sequence (parity stepping; the reference has no equivalent — it is what lets
 * tests compare kernel by kernel from identical device state). Synchronous. */
int dvp_run_stage(dvp_ctx* ctx, int stage, int iter);

/* Replaces: the D2H block at the end of RunPatchMatch (APD.cu:4525-4530) + GetPlaneHypothesis /
 * GetPixelStates / GetSelectedViews / GetRadiusMap (APD.cpp:1706-1732).  Any pointer may be NULL. */
int dvp_download(dvp_ctx* ctx, float* planes, uint8_t* weak_info, uint32_t* selected_views, int32_t* radius);

/* Raw buffer access by id (sizes in bytes must match exactly; query with dvp_buffer_bytes).  The compact WEAK-pixel
 * lists are built by dvp_upload: a DVP_BUF_WEAK written afterwards may demote WEAK pixels (as K2 / K5 do) but pixels
 * it newly marks WEAK are not picked up by K4, K9, K10 and K11 until the next upload. */
size_t dvp_buffer_bytes(dvp_ctx* ctx, int buffer);
int dvp_get_buffer(dvp_ctx* ctx, int buffer, void* dst, size_t bytes);
int dvp_set_buffer(dvp_ctx* ctx, int buffer, const void* src, size_t bytes);

/* Device time (CUDA events on the context stream) of the last dvp_run: total and per stage
 * (stages launched several times are summed). `per_stage` has DVP_STAGE_COUNT entries or is NULL.
 * Also reports how many kernels that run launched. */
int dvp_last_run_times(dvp_ctx* ctx, float* total_ms, float* per_stage_ms, int* launches);

/* Device pointers of the resident maps, for callers that keep data on the GPU between passes
 * (bench.py's HBM-resident leg; "next" row N2).  `images`/`depths`/`planes` follow dvp_inputs layouts. */
int dvp_upload_device(dvp_ctx* ctx, const dvp_inputs* device_in, const dvp_params* params);

/* "Next" row N1 — replaces the host post-processing ProcessProblem runs on the downloaded maps right after
 * RunPatchMatch (main.cpp:297-363; Connect / Label_Seek / Label_Update, APD.cpp:138-346), in place on the
 * resident maps: (1) depths outside [depth_min, depth_max] become 0 and their state UNKNOWN
 * (main.cpp:300-306); (2) for every source view, 4-connected regions of pixels that do not select the view and
 * hold fewer than 20 * (8 / scale_size)^2 pixels get the view's bit set (main.cpp:323-363).
 * `scale_size` is Problem::scale_size (8, 4, 2 or 1).  Follow with dvp_download. Synchronous.
 * `device_ms` (may be NULL) receives the device time of the call. */
int dvp_restore_visibility(dvp_ctx* ctx, int scale_size, float* device_ms);

/* ---- "Next" row N2: the multi-scale schedule with every map resident in HBM between passes ------------------------
 * Replaces the file round trips between the (view, pass) jobs of main() (main.cpp:449-511): ProcessProblem's writes
 * (main.cpp:365-376), InuputInitialization's reads + host rescales (APD.cpp:1147-1180, 1427-1456) and
 * SupportInitialization's (APD.cpp:1615-1668).  Pyramid level 0 is the coarsest (scale_size 2^num_levels ... 2; the
 * reference never runs scale 1).  Per level the caller supplies what the reference computes with OpenCV or reads from
 * other tools: the resized grey image, the edge map and the label map; and for level 0 the FIRST_INIT plane prior.
 * All views of a scene must share one full-resolution size.  Views run in index order (the reference's order). */
typedef struct dvp_scene dvp_scene;
/* device < 0 creates a host-only scene: it answers dvp_scene_level_size / dvp_scene_pass_params and refuses everything
 * that needs a GPU (DVP_ERR_STATE) — lets the schedule be checked where no GPU is present. */
dvp_scene* dvp_scene_create(int device, int num_views, int num_levels);
void dvp_scene_destroy(dvp_scene* scene);
/* Level size as InuputInitialization computes it: round(full * (1 / scale)) (APD.cpp:1119-1123). */
int dvp_scene_level_size(dvp_scene* scene, int full_w, int full_h, int level, int* w, int* h);
/* PatchMatchParams of pass 0 (FIRST_INIT / REFINE_INIT) or 1..3 (REFINE_ITER) of round `level` (main.cpp:452-505);
 * depth_min / depth_max / num_images are per view and left at their defaults. */
int dvp_scene_pass_params(dvp_scene* scene, int level, int pass, dvp_params* out);
/* PatchMatch iterations per pass: 3 in the reference (main.cpp:481, 502).  0 leaves only the race-free stages, which
 * makes whole-schedule results reproducible bit for bit (used by the chain parity test). */
int dvp_scene_set_max_iterations(dvp_scene* scene, int iterations);
/* Full-resolution camera (cam.txt), size and source views (pair.txt) of one view. */
int dvp_scene_set_view(dvp_scene* scene, int view, const dvp_camera* cam_full, int full_w, int full_h, int num_src, const int* src_views);
/* Per-level data of one view, host or device pointers: image [h][w] f32 (required), edge u8 / label i32 (or NULL). */
int dvp_scene_set_level(dvp_scene* scene, int view, int level, const float* image, const uint8_t* edge, const int32_t* label);
/* Row N2, image pyramid: every level image of one view from its full-resolution grey image [full_h][full_w] u8 (what
 * cv::imread(.., IMREAD_GRAYSCALE) returns; host or device), exactly as InuputInitialization builds them — convertTo(CV_32FC1),
 * then cv::resize of the FULL image to each level's size (APD.cpp:1057-1060, 1119-1132; dvp_resize_linear_f32).
 * compute_priors: bit 0 = every level's edge map (dvp_scene_compute_edges), bit 1 = every level's label map
 * (dvp_label_segment of the full image with the level's scale) — together what GetProblemEdges prepares (main.cpp:193-246).
 * With both bits a view needs nothing but its image, camera and source list.  dvp_scene_set_label overrides a level's
 * label map (NULL = none); dvp_scene_get_label copies it out. */
int dvp_scene_set_image(dvp_scene* scene, int view, const uint8_t* image, int compute_priors);
int dvp_scene_get_label(dvp_scene* scene, int view, int level, int32_t* label);
int dvp_scene_set_label(dvp_scene* scene, int view, int level, const int32_t* label);
int dvp_scene_get_image(dvp_scene* scene, int view, int level, float* image);
/* FIRST_INIT prior of one view at level 0: [h0][w0][4] (world normal, depth), APD.cpp:1410-1420. */
int dvp_scene_set_initial_planes(dvp_scene* scene, int view, const float* planes);
/* One pass over every view (one inner loop of main.cpp:452-511); view v runs with seed + v. */
int dvp_scene_run_pass(dvp_scene* scene, int level, int pass, uint64_t seed);
/* One (view, pass) job — one ProcessProblem — for callers that deal views to several GPUs. */
int dvp_scene_run_view(dvp_scene* scene, int view, int level, int pass, uint64_t seed);
/* The exchange step of a farmed pass: a view's depth map is all another rank needs from it (geometric consistency reads
 * the source views' depths, APD.cpp:1147-1166).  dvp_scene_depth_map returns the device buffer of a view this rank
 * owns (to broadcast from); dvp_scene_remote_depth sizes and returns the buffer of a view owned elsewhere (to
 * receive into) and marks it usable as a source. */
int dvp_scene_depth_map(dvp_scene* scene, int view, float** device_ptr, int* w, int* h);
int dvp_scene_remote_depth(dvp_scene* scene, int view, int w, int h, float** device_ptr);
/* The whole schedule: num_levels rounds x 4 passes x all views. `device_ms` (may be NULL): summed device time. */
int dvp_scene_run(dvp_scene* scene, uint64_t seed, float* device_ms);
/* Current maps of one view (size returned in *w, *h); destinations may be host or device memory, or NULL. */
int dvp_scene_get_view(dvp_scene* scene, int view, int* w, int* h, float* planes, uint8_t* weak_info, uint32_t* selected_views, int32_t* radius);
int dvp_scene_stats(dvp_scene* scene, double* device_ms, long long* passes);
/* RescaleMatToTargetSize (APD.cpp:1773-1796, swapped scale factors reproduced) on device memory; elem_bytes 1, 4 or 16. */
int dvp_rescale_map(int device, const void* src, int src_w, int src_h, void* dst, int dst_w, int dst_h, int elem_bytes);

/* ---- the per-view farm inside the library -----------------------------------------------------------------
 * One resident scene and one host thread per GPU of the box; the views of every pass are dealt round robin (view v runs on
 * devices[v % num_devices]), the data path has no collective.  The one exchange step per pass — a view's fresh depth map
 * goes to the GPUs that own a view listing it as a source — is a peer copy (NVLink) straight between the scenes' device
 * buffers.  The reference picks ONE GPU by argv[2] and loops over all views in one process (main.cpp:430-434, 452-507).
 * Within a pass a GPU sees its own views' fresh maps and the other GPUs' previous-pass maps (block Gauss-Seidel); with one
 * device this is dvp_scene_run exactly.  The setters mirror the dvp_scene_* ones and replicate their input on every GPU. */
typedef struct dvp_farm dvp_farm;
dvp_farm* dvp_farm_create(int num_devices, const int* devices, int num_views, int num_levels);
void dvp_farm_destroy(dvp_farm* farm);
int dvp_farm_num_devices(dvp_farm* farm);
int dvp_farm_owner(dvp_farm* farm, int view);     /* the CUDA device that runs this view */
int dvp_farm_set_max_iterations(dvp_farm* farm, int iterations);
int dvp_farm_set_view(dvp_farm* farm, int view, const dvp_camera* cam_full, int full_w, int full_h, int num_src, const int* src_views);
int dvp_farm_set_level(dvp_farm* farm, int view, int level, const float* image, const uint8_t* edge, const int32_t* label);
int dvp_farm_set_image(dvp_farm* farm, int view, const uint8_t* image, int compute_priors);
int dvp_farm_compute_edges(dvp_farm* farm, int view, int level);
int dvp_farm_set_initial_planes(dvp_farm* farm, int view, const float* planes);
/* The whole schedule (main.cpp:449-511).  wall_ms: host wall clock of the call; exchange_ms: what the slowest GPU thread
 * spent in the exchange steps (peer copies and their two barriers).  Either may be NULL. */
int dvp_farm_run(dvp_farm* farm, uint64_t seed, float* wall_ms, float* exchange_ms);
long long dvp_farm_exchange_bytes(dvp_farm* farm);   /* bytes the last dvp_farm_run moved between GPUs */
int dvp_farm_get_view(dvp_farm* farm, int view, int* w, int* h, float* planes, uint8_t* weak_info, uint32_t* selected_views, int32_t* radius);

/* dvp_upload whose large maps (plane hypotheses, the 1+S images, the depth maps: ~3/4 of the bytes) travel on a second
 * stream while the next dvp_run already executes K1..K3 and K5; K4 waits for the planes, K6 for everything.  Host
 * buffers must be page-locked for the copies to be asynchronous and must stay valid and unchanged until a synchronising
 * call returns (dvp_run with sync != 0, dvp_run_stage, dvp_download, dvp_get_buffer).  Results are those of dvp_upload. */
int dvp_upload_overlapped(dvp_ctx* ctx, const dvp_inputs* in, const dvp_params* params);

/* ---- "Next" row N3: depth-map fusion on the device -------------------------------------------------------------------
 * Replaces RunFusion, the "ETH version" main() calls (APD.cpp:1809-1960, main.cpp:514): every pixel of every view is
 * lifted to 3-D, projected into the view's sources, checked for reprojection / depth / normal consistency, and — if
 * accepted — emitted as one coloured point while the source pixels that agreed are masked so that they do not emit
 * the same surface point again.  The reference does this in one sequential loop (views in order, pixels in raster
 * order) and its result depends on that order; the device path reproduces exactly that order's result (deterministic
 * reservations: a pixel decides once no earlier undecided pixel claims one of its source cells).
 * One view = what the reference holds after its loading loop (APD.cpp:1841-1873): camera and colour image already
 * rescaled to the depth map's size (RescaleImageAndCamera), weak map rescaled to it (RescaleMatToTargetSize). */
typedef struct dvp_fusion_view {
	dvp_camera camera;        /* K, R, t at the depth map's size                                                   */
	int32_t width, height;    /* of the depth map                                                                  */
	const float* depth;       /* [h][w]     depths.dmb                                                             */
	const float* normal;      /* [h][w][3]  APD_normals.dmb (world frame)                                          */
	const uint8_t* image;     /* [h][w][3]  colour image, channel order kept as given (BGR in the reference)       */
	const uint8_t* weak;      /* [h][w]     DVP_WEAK / DVP_STRONG / DVP_UNKNOWN (weak.bin); may be NULL in modes 1, 2 */
	const uint8_t* block;     /* [h][w] or NULL: pixels < 128 are not fused (the optional blocks/ folder)          */
	int32_t num_src;          /* Problem::src_image_ids.size(), <= 32                                              */
	const int32_t* src_views; /* [num_src] view indices (imageIdToindexMap applied)                                */
} dvp_fusion_view;

typedef struct dvp_fusion dvp_fusion;
dvp_fusion* dvp_fusion_create(int device, int num_views);
void dvp_fusion_destroy(dvp_fusion* f);
/* Copies one view's maps into the fusion object (host or device pointers). */
int dvp_fusion_set_view(dvp_fusion* f, int view, const dvp_fusion_view* v);
/* Same, with depth and normal taken from a [h][w][4] (world normal, depth) plane map as dvp_download / dvp_scene_get_view
 * return it (host or device pointer) — the split ProcessProblem does before writing depths.dmb / APD_normals.dmb
 * (main.cpp:300-306).  v->depth and v->normal are ignored. */
int dvp_fusion_set_view_planes(dvp_fusion* f, int view, const dvp_fusion_view* v, const float* planes);
/* N2 -> N3 hand-off: registers every view of a scene with a fusion object (same device) straight from the scene's
 * device buffers — what the reference passes through depths.dmb / APD_normals.dmb / weak.bin (main.cpp:365-370 ->
 * APD.cpp:1851-1858) — with the camera rescaled to the map size (RescaleImageAndCamera, APD.cpp:1750-1771) and the
 * scene's source lists.  images[v]: [h][w][3] u8 colour image of view v at its current map size (host or device; the
 * reference's cv::resize of the jpg); blocks (or blocks[v]) may be NULL. */
int dvp_scene_fuse_views(dvp_scene* scene, dvp_fusion* f, const uint8_t* const* images, const uint8_t* const* blocks);
/* Which of the reference's three fusion routines runs: 0 RunFusion (ETH, the one main() calls; default),
 * 1 RunFusion_TAT_Intermediate (APD.cpp:1962-2130), 2 RunFusion_TAT_advanced (APD.cpp:2132-2279).  Modes 1 and 2 need no
 * weak map (v->weak may be NULL) and reproduce the reference's per-view `diff` vector, which keeps a source's measures
 * from the last pixel that evaluated it (APD.cpp:2052, 2216). */
int dvp_fusion_set_mode(dvp_fusion* f, int mode);
/* Clears the fusion masks (APD.cpp:1867) and the point list. */
int dvp_fusion_reset(dvp_fusion* f);
/* One iteration of the reference's outer loop (APD.cpp:1875-1957): fuses view `view` against the current masks and
 * appends its points.  `device_ms` may be NULL. */
int dvp_fusion_run_view(dvp_fusion* f, int view, float* device_ms);
/* dvp_fusion_reset + every view in index order. */
int dvp_fusion_run(dvp_fusion* f, long long* num_points, float* device_ms);
long long dvp_fusion_num_points(dvp_fusion* f);
/* Points [first, first + count) in the reference's order: 6 floats each (coord xyz, colour in the image's channel
 * order — PointList, main.h:69-72). */
int dvp_fusion_get_points(dvp_fusion* f, float* dst, long long first, long long count);
int dvp_fusion_get_mask(dvp_fusion* f, int view, uint8_t* dst);
/* Stage stepping for parity tests, about the last dvp_fusion_run_view: the mask-independent candidates
 * (cells / terms: [h*w][num_src], -1 = source not consistent), the per-pixel decision (used: bit j = source j
 * contributed, 0 = no point) and the number of reservation rounds it took.  Any pointer may be NULL. */
int dvp_fusion_last_view(dvp_fusion* f, int32_t* cells, float* terms, uint32_t* used, int* rounds);
/* Index of the view dvp_fusion_last_view reports on (its buffers are sized by THAT view's h, w and num_src), or
 * DVP_ERR_STATE when no view has run since the last reset. */
int dvp_fusion_last_view_index(dvp_fusion* f);
/* ExportPointCloud (APD.cpp:842-882): binary little-endian PLY, x y z float + 3 uchar colours. */
int dvp_fusion_write_ply(dvp_fusion* f, const char* path);

/* ---- "Next" row N4, edge half: the depth-edge prior ------------------------------------------------------------------
 * Replaces EdgeSegment(scale, image, mode 0, use_canny = true) (APD.cpp:348-466) as GetProblemEdges calls it per view
 * and pyramid level (main.cpp:193-226); its result is the `edge` input of dvp_upload / dvp_scene_set_level.
 * image: [height][width] u8 grey levels (the level image after convertTo(CV_8UC1)), host or device; edge: [height][width]
 * u8, 0 / 255, host or device.  thresholds (may be NULL): the two Canny thresholds derived from the histogram median
 * (APD.cpp:405-432).  Bit-exact with OpenCV's Canny (3x3 Sobel, L2 gradient) followed by the reference's border
 * clean-up.  width, height >= 3.  Stateless and synchronous. */
int dvp_edge_segment(int device, const uint8_t* image, int width, int height, uint8_t* edge, int32_t* thresholds, float* device_ms);
/* The same for a level of a scene, from the level image given to dvp_scene_set_level (rounded to 8 bits as
 * main.cpp:209 does): the edge map stays in HBM as that level's edge input.  dvp_scene_get_edges copies it out. */
int dvp_scene_compute_edges(dvp_scene* scene, int view, int level);
int dvp_scene_get_edges(dvp_scene* scene, int view, int level, uint8_t* edge);

/* ---- Row N4, label half: the region-label prior ----------------------------------------------------------------------------
 * Replaces EdgeSegment(scale, image_uint, 1) (APD.cpp:348-402, 437-499) as GetProblemEdges calls it on the FULL-resolution
 * 8-bit grey image (main.cpp:229-241; `scale` = log2 of Problem::scale_size); its result is the `label` input of
 * dvp_upload / dvp_scene_set_label: 0 on region boundaries, -1 in regions of at most weak_tex_num pixels, a positive id
 * elsewhere.  The hot path only compares labels for equality and tests `> 0`, `== 0`, `== -1` (APD.cu:3461, 3629,
 * 3857-3886), so the ids are this library's own (root pixel + 1): the PARTITION and the classes are the reference's.
 * labels: [new_rows][new_cols] int32 with (new_cols, new_rows) = dvp_label_size(cols, rows, scale); edge_small (may be
 * NULL): the (rows/2)/2 x (cols/2)/2 edge image after the Hough lines were drawn, for stage-wise checks.  Host or device
 * pointers.  cols, rows >= 16.  Stateless and synchronous. */
int dvp_label_size(int cols, int rows, int scale, int* new_cols, int* new_rows);
int dvp_label_segment(int device, const uint8_t* image, int cols, int rows, int scale, int32_t* labels, uint8_t* edge_small, float* device_ms);

/* ---- Row N2, image pyramid: the level image every pass of a view reads --------------------------------------------------
 * Replaces cv::resize(image, scaled, Size(new_cols, new_rows), 0, 0, INTER_LINEAR) of the float grey image in
 * InuputInitialization (APD.cpp:1119-1140) and GetProblemEdges (main.cpp:203-209).  cv::resize is OpenCV's; what is
 * reproduced is its generic bilinear path for CV_32F, bit-exact with OpenCV 4.13 when its IPP back end is off (an
 * IPP-enabled build differs by up to 0.015 grey levels).  src [src_h][src_w], dst [dst_h][dst_w], host or device. */
int dvp_resize_linear_f32(int device, const float* src, int src_w, int src_h, float* dst, int dst_w, int dst_h);

/* ---- Row N4, second half: the reference's on-disk exchange formats (host code, no GPU) ------------------------------
 * .bin / .dmb files of WriteBinMat / ReadBinMat (APD.cpp:548-573, 630-648): int32 version = 1, rows, cols, OpenCV type
 * code (CV_8U 0, CV_32S 4, CV_32F 5, CV_32FC3 21, ...), then rows * cols * elemSize bytes.  DVP_ERR_STATE: cannot open /
 * short file; DVP_ERR_UNSUPPORTED: version != 1 (the reference's "Version error"). */
int dvp_io_binmat_header(const char* path, int32_t* rows, int32_t* cols, int32_t* cv_type);
int dvp_io_read_binmat(const char* path, void* data, size_t bytes);          /* bytes must equal the payload size */
int dvp_io_write_binmat(const char* path, int32_t rows, int32_t cols, int32_t cv_type, const void* data);
/* writeDepthDmb (channels 1) / writeNormalDmb (channels 3), APD.cpp:575-628: int32 1, h, w, channels, then floats. */
int dvp_io_write_dmb(const char* path, int32_t rows, int32_t cols, int32_t channels, const float* data);
/* ReadCamera (APD.cpp:651-692), the TAT & ETH cam.txt layout; the centre is computed as the reference does. */
int dvp_io_read_camera(const char* path, dvp_camera* cam);
/* GenerateSampleList (main.cpp:127-170): pair.txt.  With ref_ids == NULL only *num_views is returned.  src_ids is
 * [max_views][DVP_MAX_IMAGES]; sources with score <= 0 are dropped as in the reference.  DVP_ERR_UNSUPPORTED: a negative view
 * count, or a view with more than DVP_MAX_IMAGES - 1 usable sources (more than a context can take; never silently cut). */
int dvp_io_read_pairs(const char* path, int32_t max_views, int32_t* num_views, int32_t* ref_ids, int32_t* num_src, int32_t* src_ids);

/* ---- Parity instrumentation for the racy stage (tests only; dvp_run never takes this path) --------------------------
 * The reference's strong sweep reads, in diagonal direction 4 only, cost and plane of pixels of the colour it is writing
 * in the same launch (APD.cu:2039, 2071-2074: the colour fix covers `dir_index > 4`) — so what a pixel ends up with
 * depends on which of those pixels were already rewritten when it looked.  Whatever the timing, direction 4 contributes
 * ONE candidate: some ladder pixel at offset m along the diagonal, i.e. (x - 5 - m, y - 5 - m), whose plane is read three
 * times — for scoring (APD.cu:2084 / 2133), for the depth test and for the copy at acceptance (APD.cu:2559-2563) — each
 * time before or after that pixel's own update, or torn between the two (the reference build loads a float4 plane with
 * four 32-bit loads while its owner replaces it with one 128-bit store).
 * dvp_debug_race_explain answers, for an OBSERVED result of K7 (red = 0) / K8 (red = 1) launched from the state this
 * context currently holds: which pixels are reproduced, in all five output buffers bit for bit, by the production
 * arithmetic under SOME such choice?  Phase 1 tries every offset in `offsets` with each read entirely before / after
 * (8 combinations) on the whole image; with `tear` != 0 phase 2 tries, on the pixels phase 1 left over, the remaining 4088
 * component mixtures per offset and the scoring read repeated per source view (some views before, the others after).
 * The state is not modified (results go to shadow buffers).
 * planes_before / planes_after: [H][W][4] plane maps before the launch and after it (the observed result's);
 * exp_*: the observed result — planes [H][W][4], costs [H][W], selected [H][W], view_weight [H][W][32], rand [H][W][6];
 * explained: [H][W] out, 1 = reproduced (pixels the launch does not process: 1 iff unchanged);
 * stats (may be NULL): {unexplained after phase 1, unexplained after phase 2, forced launches}.  Host or device pointers. */
int dvp_debug_race_explain(dvp_ctx* ctx, int iter, int red, const int32_t* offsets, int num_offsets, const float* planes_before, const float* planes_after,
                           const float* exp_planes, const float* exp_costs, const uint32_t* exp_selected, const uint8_t* exp_view_weight, const uint32_t* exp_rand,
                           int tear, uint8_t* explained, long long* stats);
/* Measurement instrumentation: texture fetches issued on this context since the last reset (every fetch site of the
 * NCC / reprojection code tallies itself).  Only the instrumented build of the same sources (libdvp_mvs_count.so, `make
 * count` in csrc/) counts; the first call arms the counter and returns 0; the production library returns
 * DVP_ERR_UNSUPPORTED.  bench.py uses it, outside the timed region, to report measured fetches per launch. */
long long dvp_debug_fetch_count(dvp_ctx* ctx, int reset);
int dvp_debug_sweep_forced_d4(dvp_ctx* ctx, int iter, int red, int m, int ncc_from_after, int accept_from_after);

int dvp_weak_count(dvp_ctx* ctx);
int dvp_last_cuda_error(dvp_ctx* ctx);
void* dvp_stream(dvp_ctx* ctx); /* cudaStream_t */

#ifdef __cplusplus
}
#endif
#endif /* DVP_MVS_H */
