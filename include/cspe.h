/*
 * cspe.h — C ABI of libcspe.so: the B200-native (sm_100a) per-frame annotation hot path
 * behind xander683/ConstructionScenePoseEstimation's generate_construction_data.py
 * (abbreviated gcd.py below; all gcd.py:line citations are into that file of the reference).
 *
 * The reference has no FFI / plugin API (it is one Python script); its de-facto boundary is
 * the Replicator annotator-dict surface consumed at gcd.py:1669-1681 (RGB, depth),
 * gcd.py:1780-1790 / 1916-1922 (bounding_box_3d), gcd.py:1818-1842 (instance_segmentation)
 * and the label file written at gcd.py:2055-2072.  Every entry point below names the span of
 * gcd.py it replaces (or, for stages the reference leaves as a hole, the span where it plugs
 * in).  The Python host side (constructionsceneposeestimation_b200.writer) binds these with
 * ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns int: 0 = CSPE_OK, < 0 = CSPE_ERR_*; cspe_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread;
 *   - all data pointers are DEVICE pointers unless the parameter name ends in _host;
 *   - the caller owns every buffer; the library never allocates user-visible memory;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*) and never
 *     synchronises, so calls are CUDA-graph capturable;
 *   - re-entrant per stream; no global state besides the thread-local error string and a
 *     per-device cache of immutable device attributes;
 *   - stream semantics: the kernels are launched with programmatic stream serialisation, but every
 *     plain entry point waits for the previous kernel in `stream` before it reads or writes memory,
 *     i.e. it behaves like an ordinary stream-ordered launch.  The `*_overlapped` variants relax
 *     that for steady-state pipelines (each states what the caller must guarantee in return).
 */
#ifndef CSPE_H_
#define CSPE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSPE_ABI_VERSION 1

enum {
  CSPE_OK = 0,
  CSPE_ERR_INVALID_ARGUMENT = -1, /* null pointer, negative size, misaligned buffer */
  CSPE_ERR_UNSUPPORTED = -2,      /* shape outside what the kernels were built for */
  CSPE_ERR_CUDA = -3,             /* a CUDA runtime call or launch failed */
  CSPE_ERR_NO_DEVICE = -4         /* no sm_100 device is current */
};

/* ---- fixed layouts -------------------------------------------------------------------- */

/* per-instance scan result, int32[5]: pixel count, then INCLUSIVE tight box.
 * Absent instance: {0, W, H, -1, -1}. */
#define CSPE_SCAN_FIELDS 5
enum { CSPE_SCAN_COUNT = 0, CSPE_SCAN_XMIN = 1, CSPE_SCAN_YMIN = 2, CSPE_SCAN_XMAX = 3, CSPE_SCAN_YMAX = 4 };

/* Replicator bounding_box_3d record as the reference indexes it (gcd.py:562-564:
 * [0] semanticId u32, [1..6] x_min,y_min,z_min,x_max,y_max,z_max f32, [7] transform f32[4][4]
 * (USD row-vector convention), [8] occlusionRatio f32) = 96 bytes. */
#define CSPE_BBOX3D_RECORD_BYTES 96

/* per-frame camera block, double[CSPE_CAM_STRIDE]:
 *  [0..2]  t    camera position in world            (gcd.py:599 ExtractTranslation)
 *  [3..11] Rcw  camera->world rotation, row-major, column-vector convention
 *               (gcd.py:600-601 ExtractRotationMatrix().GetTranspose()), USD camera axes
 *               (-Z forward, +Y up)
 *  [12] fx [13] fy [14] cx [15] cy                   (gcd.py:646-649)
 *  [16] near clip  [17] far clip                     (gcd.py:1437)
 *  [18] W  [19] H  [20..23] reserved (0)                                                  */
#define CSPE_CAM_STRIDE 24

/* per-object pose block, double[CSPE_POSE_STRIDE]:
 *  [0..2] t_co   object centre in the camera frame
 *  [3..6] q_co   object->camera rotation, quaternion xyzw, w >= 0
 *  [7..9] center_world  [10..12] size_world  [13..15] euler xyz degrees (gcd.py:553-584) */
#define CSPE_POSE_STRIDE 16

/* object flags (uint8) */
enum {
  CSPE_OBJ_HAS_RECORD = 1,  /* a bbox3d record was resolved for the slot */
  CSPE_OBJ_ANY_FRONT = 2,   /* at least one 3D-box corner has z > near */
  CSPE_OBJ_ALL_FRONT = 4,   /* all eight corners have z > near */
  CSPE_OBJ_POSE_VALID = 8,  /* rotation part is finite with det > 0 (scipy would not raise) */
  CSPE_OBJ_APPROX_RECORD = 16 /* the 3D box / pose come from a record that only approximates the object: a mesh
                               * record standing in for an object whose root prim has none (the reference reads
                               * the whole object's bound from the live USD stage there, gcd.py:1977-2023) */
};
/* obj_record entries (K2 input): OR this bit into a record index to mark it as such a stand-in */
#define CSPE_OBJ_RECORD_APPROX_BIT (1 << 30)

/* keypoint visibility, COCO convention */
enum { CSPE_KP_OUT = 0, CSPE_KP_OCCLUDED = 1, CSPE_KP_VISIBLE = 2 };

/* One emitted label record (array-of-structs so a host can view the D2H buffer with one
 * numpy structured dtype).  408 bytes, 8-byte aligned. */
typedef struct cspe_record {
  int32_t frame;       /* global frame id (frame_base + batch index) */
  int32_t inst_idx;    /* slot = reference inst_idx (gcd.py:1876-1886) */
  int32_t class_id;    /* gcd.py:69-106 */
  int32_t count;       /* visible pixels */
  int32_t x_min, y_min, x_max, y_max; /* tight box, inclusive */
  int32_t flags;       /* CSPE_OBJ_* */
  int32_t loose[4];    /* projected 3D box clipped to the image, inclusive ints; empty -> {0,0,-1,-1} */
  int32_t pad0;
  float occlusion;     /* 1 - visible_frac */
  float fill;          /* count / tight_area */
  float truncation;    /* 1 - clipped_area / unclipped_area of the projected 3D box */
  float visible_frac;  /* min(1, count / loose_area) */
  float yolo[4];       /* cx/W, cy/H, w/W, h/H of the tight box */
  double uv[16];       /* 8 projected corners (u,v) in pixels */
  double z[8];         /* camera depth of each corner (+ = in front) */
  double pose[CSPE_POSE_STRIDE];
} cspe_record;

#define CSPE_NUM_CLASSES 10 /* class ids 0..9, gcd.py:69-106 */

/* depth statistics block per frame (gcd.py:314-359), see cspe_depth_stats */
typedef struct cspe_depth_stats_t {
  int64_t valid_pixels; /* isfinite & > 0 */
  int64_t zero_pixels;  /* == 0 */
  int64_t inf_pixels;   /* isinf */
  int64_t total_pixels;
  float depth_min;      /* over valid pixels; 0 if none */
  float depth_max;
  double depth_sum;     /* f64 sum over valid pixels (mean = sum / valid_pixels) */
} cspe_depth_stats_t;

/* ---- library ---------------------------------------------------------------------------- */

int cspe_version(void);              /* CSPE_ABI_VERSION */
const char* cspe_last_error(void);   /* thread-local; "" if none */
/* SM count and compute capability of the current device. */
int cspe_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- K1: instance-ID mask scan  (fills the hole at gcd.py:1908-1910 / 2066-2069; input is
 * the instance_segmentation annotator attached at gcd.py:1475) ----------------------------
 * mask     uint32 [B][H][W] row-major, Replicator native ids
 * id2slot  int32 LUT: slot = id < lut_len ? id2slot[frame*lut_stride + id] : -1;
 *          slot < 0 or >= N means "ignore" (BACKGROUND 0, UNLABELLED 1, unmapped meshes).
 *          Several ids may map to one slot (mesh -> object aggregation, gcd.py:1858-1891):
 *          counts add, boxes union.  lut_stride = 0 shares one LUT across the batch.
 * out      int32 [B][N][5]  (CSPE_SCAN_*), fully overwritten.
 * Single pass over the mask: 4*H*W bytes read per frame.                                  */
int cspe_mask_scan(const uint32_t* mask, int B, int H, int W,
                   const int32_t* id2slot, int lut_len, int64_t lut_stride,
                   int N, int32_t* out, void* stream);

/* K1 without the initialisation: merges this mask's pixels into an `out` that already holds
 * valid scan entries (counts add, boxes union) — e.g. a frame delivered in several tiles. */
int cspe_mask_scan_accumulate(const uint32_t* mask, int B, int H, int W,
                              const int32_t* id2slot, int lut_len, int64_t lut_stride,
                              int N, int32_t* out, void* stream);

/* K1 (accumulate) for steady-state pipelines: may START streaming the mask while the kernel
 * queued before it in `stream` is still running (programmatic dependent launch; cspe_emit
 * releases its dependents early) and only waits for it before its first merge into `out`.
 * Contract: of this call's buffers, only `out` may be read or written by that previous kernel
 * (K4 of the previous batch reading and resetting the table) — mask and id2slot must be stable. */
int cspe_mask_scan_accumulate_overlapped(const uint32_t* mask, int B, int H, int W,
                                         const int32_t* id2slot, int lut_len, int64_t lut_stride,
                                         int N, int32_t* out, void* stream);

/* 1 when cspe_mask_scan* on a [B][H][W] batch launches its full persistent grid (two CTAs on every SM).
 * A pipeline may chain the `*_overlapped` entry points across batches with double-buffered K2 / K3
 * outputs only then (see DESIGN.md "Step pipeline"); smaller batches must use the plain entry points. */
int cspe_mask_scan_fills_device(int B, int H, int W);

/* K1 plus the depth-quality statistics of gcd.py:314-359 in one call (two HBM-bound launches on
 * `stream`; a fused kernel measured slower).
 * depth float32 [B][H][W]; stats cspe_depth_stats_t[B], fully overwritten. */
int cspe_mask_scan_depth_stats(const uint32_t* mask, const float* depth, int B, int H, int W,
                               const int32_t* id2slot, int lut_len, int64_t lut_stride,
                               int N, int32_t* out, cspe_depth_stats_t* stats, void* stream);

/* ---- K2: per-object transform, projection and pose  (replaces the per-object loop
 * gcd.py:1924-1950 -> bboxDict_to_transform gcd.py:553-584, and adds corner projection and
 * the object-in-camera pose the reference leaves out) --------------------------------------
 * records     bbox3d records, frame f record r at records + (f*recs_per_frame + r)*rec_stride
 * obj_record  int32 [B][N]: record index of slot n in frame f, or -1 (no record -> flags 0); bit 30
 *             (CSPE_OBJ_RECORD_APPROX_BIT) set = stand-in record -> CSPE_OBJ_APPROX_RECORD in flags
 * cam         double [B][CSPE_CAM_STRIDE]
 * uv          double [B][N][8][2]   z double [B][N][8]   pose double [B][N][CSPE_POSE_STRIDE]
 * loose       double [B][N][4] = u_min,v_min,u_max,v_max over in-front corners (NaN if none)
 * flags       uint8  [B][N]                                                                */
int cspe_project_objects(const void* records, int rec_stride, int recs_per_frame,
                         const int32_t* obj_record, const double* cam, int B, int N,
                         double* uv, double* z, double* pose, double* loose, uint8_t* flags,
                         void* stream);

/* K2 for pipelines: identical results, but the grid may START while the kernel queued before it
 * in `stream` is still running, provided that kernel releases its dependents early (the mask
 * scan does: programmatic dependent launch); it still COMPLETES after it, so whatever is queued
 * next sees both done.  Contract: none of this call's inputs is written by that previous kernel.
 * (cspe_project_objects itself makes no such assumption.) */
int cspe_project_objects_overlapped(const void* records, int rec_stride, int recs_per_frame,
                                    const int32_t* obj_record, const double* cam, int B, int N,
                                    double* uv, double* z, double* pose, double* loose, uint8_t* flags,
                                    void* stream);

/* Object-level records for multi-mesh objects whose root prim has no bbox3d record of its own.  The reference
 * reads that object's bound from the live USD stage — BBoxCache.ComputeWorldBound(prim).ComputeAlignedRange(),
 * gcd.py:2000-2009 — which is the world-axis-aligned range of everything under the prim; here the same range is the
 * min / max over the eight world-space corners of every mesh record of the object.
 * records   the frame's record array as cspe_project_objects takes it, with room for U more records per frame:
 *           frame f, union object u is WRITTEN at records + (f*recs_per_frame + base + u)*rec_stride as a record with
 *           an identity transform whose extents are the range (float32); members are read from indices < base
 * offsets   int32 [rows][U+1], members int32 [rows][M]: object u unites records members[offsets[u] .. offsets[u+1]);
 *           rows = B with strides U+1 / M, or 1 with stride 0 (one scene for the whole batch)
 * An object none of whose members has a finite corner gets NaN extents (K2 then clears CSPE_OBJ_POSE_VALID).
 * Ordinary stream-ordered launch: queue it before the mask scan of a PDL chain, or anywhere before K2 otherwise. */
int cspe_union_records(void* records, int rec_stride, int recs_per_frame, int base, const int32_t* offsets,
                       int64_t offsets_stride, const int32_t* members, int64_t members_stride, int B, int U,
                       void* stream);

/* ---- K3: skeleton keypoint projection + depth-buffer visibility ([SPEC]; "skelroot" is
 * only a class keyword in the reference, gcd.py:105) ----------------------------------------
 * joints  float32 [B][P][J][3] world positions (Replicator skeleton_data globalTranslations)
 * depth   float32 [B][H][W] distance_to_image_plane (inf = no hit, gcd.py:318-321)
 * kp      double [B][P][J][2] (u,v)   kz double [B][P][J] camera depth   vis uint8 [B][P][J] */
int cspe_keypoints(const float* joints, int B, int P, int J,
                   const float* depth, int H, int W, const double* cam, double tol,
                   double* kp, double* kz, uint8_t* vis, void* stream);

/* K3 for pipelines: same contract as cspe_project_objects_overlapped — may start while the kernel
 * queued before it is still running (if that kernel releases its dependents early), completes
 * after it; none of this call's inputs may be written by that kernel. */
int cspe_keypoints_overlapped(const float* joints, int B, int P, int J,
                              const float* depth, int H, int W, const double* cam, double tol,
                              double* kp, double* kz, uint8_t* vis, void* stream);

/* ---- K4: occlusion ratios, order-preserving compaction, record emission and class
 * histogram  (replaces pose_list assembly gcd.py:1938-1946 and generalises the object
 * counter gcd.py:361-372) -------------------------------------------------------------------
 * scan        int32 [B][N][5] from K1;  uv/z/pose/loose/flags from K2
 * slot_class  int32 [B][N] class id per slot (-1 = empty slot)
 * records     cspe_record [B][N]: frame f's kept records are records[f*N .. f*N+n_out[f]),
 *             in increasing inst_idx order (stable)
 * n_out       int32 [B]
 * class_hist  int64 [CSPE_NUM_CLASSES], ACCUMULATED (caller zeroes it at sweep start)      */
int cspe_emit(const int32_t* scan, const double* uv, const double* z, const double* pose,
              const double* loose, const uint8_t* flags, const int32_t* slot_class,
              int B, int N, int H, int W, int min_pixels, int frame_base,
              cspe_record* records, int32_t* n_out, int64_t* class_hist, void* stream);

/* K4 for steady-state pipelines: same records, and every scan entry is re-initialised to the
 * absent-instance identity {0, W, H, -1, -1} once read, so the next batch can go through
 * cspe_mask_scan_accumulate on the same `scan` buffer without an initialisation launch. */
int cspe_emit_reset_scan(int32_t* scan, const double* uv, const double* z, const double* pose,
                         const double* loose, const uint8_t* flags, const int32_t* slot_class,
                         int B, int N, int H, int W, int min_pixels, int frame_base,
                         cspe_record* records, int32_t* n_out, int64_t* class_hist, void* stream);

/* K4 (reset form) whose frame numbering comes from device memory: record.frame = frame_base +
 * *frame_base_dev + f.  A captured CUDA graph of several batches can then be replayed for any frame
 * range by rewriting one int32 (the per-frame loop counter of gcd.py:1540 / frame_id of gcd.py:2057). */
int cspe_emit_reset_scan_indirect(int32_t* scan, const double* uv, const double* z, const double* pose,
                                  const double* loose, const uint8_t* flags, const int32_t* slot_class,
                                  int B, int N, int H, int W, int min_pixels, int frame_base,
                                  const int32_t* frame_base_dev, cspe_record* records, int32_t* n_out,
                                  int64_t* class_hist, void* stream);

/* S6 / f3: YOLO label text on the DEVICE (the label file written at gcd.py:2055-2072, YOLO flavour):
 * for every frame f the lines "class cx cy w h\n" (six decimals, byte-identical to
 * cspe_format_yolo_host and to Python's f"{c} {cx:.6f} ...") of its n_out[f] kept records, contiguous at
 * text + f * frame_stride.  n_bytes int32 [B] = size of the frame's text (bytes past frame_stride are
 * dropped but counted; -1 = a box value that is not finite or >= 2^20).  38 bytes per record for class
 * ids 0..9, so frame_stride = 48 * N is always enough for valid boxes.  D2H then carries ~2 KB of text
 * per 1080p frame instead of 26 KB of records. */
int cspe_format_yolo(const cspe_record* records, const int32_t* n_out, int B, int N,
                     char* text, int64_t frame_stride, int32_t* n_bytes, void* stream);

/* S6 / f3: COCO annotations on the DEVICE.  For every frame f the objects
 *   {"id": i, "image_id": record.frame, "category_id": c, "bbox": [x_min, y_min, w, h], "area": count,
 *    "iscrowd": 0, "occlusion": o, "truncation": t}
 * of its n_out[f] kept records, each preceded by ", " unless it is annotation 1, contiguous at
 * text + f * frame_stride: the frames' texts concatenated in order are the bytes cspe_format_coco_host writes
 * for records without keypoints (the ratios are printed as Python prints round(x, 6)); cspe_concat_rows_host
 * does that concatenation.  Annotation ids count up across frames and calls:
 * ann_state int64[2] (device) = {annotations printed so far, internal ticket}; zero both at sweep start, the
 * kernel advances [0] by the batch total.  n_bytes int32 [B] as for cspe_format_yolo (-1 = a ratio that is not
 * finite or >= 2^20; bytes past frame_stride are dropped but counted); 224 bytes per record always suffice. */
int cspe_format_coco(const cspe_record* records, const int32_t* n_out, int B, int N, int64_t* ann_state,
                     char* text, int64_t frame_stride, int32_t* n_bytes, void* stream);

/* The strided text of cspe_format_yolo / cspe_format_coco back to back (device): packed receives row f's
 * min(max(n_bytes[f], 0), frame_stride) bytes behind the rows before it (bytes past capacity are dropped);
 * total_bytes int64[1] = the size of the whole chunk.  A host that keeps a batch's text then takes ONE slice. */
int cspe_pack_rows(const char* text, int64_t frame_stride, const int32_t* n_bytes, int B, char* packed,
                   int64_t capacity, int64_t* total_bytes, void* stream);

/* ---- plumbing for captured step graphs ----------------------------------------------------- */

/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream) through the library's runtime, so a host
 * driver can capture the D2H read-back of a batch into the same CUDA graph as the kernels. */
int cspe_memcpy_async(void* dst, const void* src, size_t bytes, void* stream);

/* Diagnostics: node count, edge count and number of PROGRAMMATIC dependency edges of a cudaGraph_t
 * (passed as void*) — a captured step graph must keep one programmatic edge per overlapped launch. */
int cspe_graph_edge_kinds(void* cuda_graph, int* num_nodes, int* num_edges, int* num_programmatic);

/* ---- next rows ------------------------------------------------------------- */

/* f1: depth -> coloured point cloud (gcd.py:616-711).  One frame per call.
 * depth float32 [H][W]; rgb uint8 [H][W][C] (C = 3 or 4) or NULL (white);
 * cam double[CSPE_CAM_STRIDE] (uses t, Rcw, fx, fy, cx, cy exactly as gcd.py:664-685 does);
 * out double [capacity][6] x,y,z,r,g,b in row-major pixel order (stable compaction);
 * n_points int64[1]; scratch: cspe_pointcloud_workspace_bytes(H, W) bytes, 16-byte aligned.
 * Valid pixel: isfinite & > 0 & < 250 (gcd.py:655).  Points beyond capacity are dropped
 * but still counted. */
size_t cspe_pointcloud_workspace_bytes(int H, int W);
int cspe_depth_to_pointcloud(const float* depth, const uint8_t* rgb, int rgb_channels,
                             int H, int W, const double* cam, double* out, int64_t capacity,
                             int64_t* n_points, void* workspace, void* stream);

/* f1 for a batch (the per-frame loop around gcd.py:1720-1759): depth float32 [B][H][W]; rgb uint8
 * [B][H][W][C] or NULL; cam double [B][CSPE_CAM_STRIDE]; the frames' points are compacted back to back
 * into out double [capacity][6], each frame in row-major pixel order; offsets int64 [B + 1] (device) =
 * first point of every frame, offsets[B] = total (points beyond capacity are dropped but counted).  The
 * "max <= 1 -> x255" colour rule (gcd.py:693) is applied per frame, as the reference applies it per call.
 * Two launches for the whole batch; scratch: cspe_pointcloud_batch_workspace_bytes(B, H, W) bytes, 16-byte
 * aligned. */
size_t cspe_pointcloud_batch_workspace_bytes(int B, int H, int W);
int cspe_depth_to_pointcloud_batch(const float* depth, const uint8_t* rgb, int rgb_channels,
                                   int B, int H, int W, const double* cam, double* out,
                                   int64_t capacity, int64_t* offsets, void* workspace, void* stream);

/* f2: depth statistics alone (gcd.py:314-359); stats cspe_depth_stats_t[B]. */
int cspe_depth_stats(const float* depth, int B, int H, int W, cspe_depth_stats_t* stats,
                     void* stream);

/* f4: depth visualisation (gcd.py:1691-1709): out uint8 [B][H][W][3] =
 * lut_bgr[((d - min) / (max - min + 1e-6) * 255) as uint8] over valid pixels, lut_bgr[0] elsewhere;
 * frames without a valid pixel come out black.  stats = output of cspe_depth_stats (device);
 * lut_bgr uint8 [256][3] (the reference's table is cv2.COLORMAP_JET). */
int cspe_depth_colormap(const float* depth, int B, int H, int W, const cspe_depth_stats_t* stats,
                        const uint8_t* lut_bgr, uint8_t* out, void* stream);

/* f4: RGB(A) -> BGR (gcd.py:1671 cv2.COLOR_RGB2BGR on rgb_image[..., :3]);
 * rgb uint8 [num_pixels][channels >= 3], bgr uint8 [num_pixels][3]. */
int cspe_rgb_to_bgr(const uint8_t* rgb, int channels, int64_t num_pixels, uint8_t* bgr, void* stream);

/* f3: "%.6f" text of a matrix, formatted on the device — the byte stream of
 *   np.savetxt(path, depth_data, delimiter=' ', fmt='%.6f')                              gcd.py:1688
 *   np.savetxt(path, xyzrgb, fmt='%.6f', delimiter=' ', header='x y z r g b', comments='') gcd.py:1752
 * values: float32 or float64 [max_rows][cols] (dtype = CSPE_DTYPE_F32 / CSPE_DTYPE_F64);
 * n_rows: device int64[1] = live rows (e.g. n_points of cspe_depth_to_pointcloud), or NULL = max_rows;
 * header: HOST string of < 63 bytes without '\n', written first followed by '\n', or NULL;
 * text: device char[capacity]; n_bytes: device int64[1] = size of the complete text (bytes past
 *   capacity are dropped but counted; -1 = a finite |value| >= 2^128 was met, text undefined);
 * split_rows > 0: split_offsets[k] (device int64[ceil(rows / split_rows)]) = byte offset of row
 *   k * split_rows, so a [B*H][W] batch of depth maps is cut into B files;
 * scratch: cspe_text_workspace_bytes(max_rows, cols) bytes, 8-byte aligned.
 * Every finite value is rounded exactly (ties to even) like printf / Python; nan, inf, -inf are
 * written as Python writes them. */
#define CSPE_DTYPE_F32 0
#define CSPE_DTYPE_F64 1
size_t cspe_text_workspace_bytes(int64_t max_rows, int cols);
int cspe_format_fixed6(const void* values, int dtype, int64_t max_rows, const int64_t* n_rows, int cols,
                       const char* header, char* text, int64_t capacity, int64_t* n_bytes,
                       int64_t split_rows, int64_t* split_offsets, void* workspace, void* stream);

/* f3: host-side YOLO serialisation of a D2H record buffer (no CUDA; all pointers are HOST
 * pointers).  Formats the first `frames` frames of records_host [B][N] / n_out_host [B] as
 * "class cx cy w h\n" lines with six decimals — byte-identical to Python's
 * f"{c} {cx:.6f} {cy:.6f} {w:.6f} {h:.6f}" — back to back into out_host; offsets_host[f] /
 * offsets_host[f+1] delimit frame f's text (offsets_host has frames + 1 entries).
 * Returns the number of bytes written, or a negative CSPE_ERR_* (buffer too small, bad n_out). */
int64_t cspe_format_yolo_host(const cspe_record* records_host, const int32_t* n_out_host, int B, int N,
                              int frames, char* out_host, int64_t capacity, int64_t* offsets_host);

/* f3: the label files of a batch, written natively (no CUDA; HOST pointers): file j is
 * "<dir>/<prefix><first_id + j, zero-padded to `digits`><suffix>" — the naming of gcd.py:2071
 * (label_%06d.json) — and holds sizes_host[j] bytes from data_host + j * stride (the layout
 * cspe_format_yolo leaves after D2H).  Existing files are truncated.  Returns the bytes written or a
 * negative CSPE_ERR_* (cspe_last_error names the file and errno). */
int64_t cspe_write_files_host(const char* dir, const char* prefix, int digits, const char* suffix,
                              int64_t first_id, int count, const char* data_host, int64_t stride,
                              const int32_t* sizes_host);

/* f3: rows of a strided host buffer back to back (no CUDA; HOST pointers): out_host receives
 * data_host[j * stride .. + sizes_host[j]) for j in [0, count) — the frames' label text of a D2H buffer as one
 * chunk.  Returns the bytes written or a negative CSPE_ERR_* (a size outside [0, stride], capacity too small). */
int64_t cspe_concat_rows_host(const char* data_host, int64_t stride, const int32_t* sizes_host, int count,
                              char* out_host, int64_t capacity);

/* f3: the "images" entries of a COCO file for frames first_id .. first_id + count (no CUDA; HOST pointers):
 *   {"id": i, "width": W, "height": H, "file_name": "rgb_%06d.png"}   joined by ", "
 * (rgb_%06d.png is the capture loop's image name, gcd.py:1672).  Returns the bytes written or a negative
 * CSPE_ERR_* (160 bytes per image always suffice). */
int64_t cspe_format_coco_images_host(int64_t first_id, int count, int width, int height, char* out_host,
                                     int64_t capacity);

/* f3: host-side label JSON of ONE frame (no CUDA; all pointers are HOST pointers) — the text
 * json.dump(label, f, indent=2, ensure_ascii=False) writes (gcd.py:608-613) for the frame dict of
 * gcd.py:2056-2064 with the object dicts of gcd.py:1938-1946 plus the added fields
 * (formats.reference_label is the Python statement of the same layout).
 * records_host[n]: the frame's kept records in inst_idx order (D2H buffer of cspe_emit);
 * camera_params_json / class_mapping_json: the already serialised values of those two keys (nested
 *   one level: continuation lines indented by two more spaces);
 * slot_strings + slot_string_offsets[2 * num_slots + 1]: per slot the JSON string literals (quotes
 *   and escapes included) of class_name, then prim_path;
 * keypoints double [num_people][num_joints][2], visibility uint8 [num_people][num_joints],
 *   person_of_slot int32 [num_slots] (-1 = no skeleton) — all three NULL when the frame has none.
 * Floats are written as Python's repr(); non-finite label values as null where the Python path maps
 * them to None.  Returns the number of bytes written or a negative CSPE_ERR_*. */
int64_t cspe_format_label_json_host(const cspe_record* records_host, int n, int64_t frame_id,
                                    const double* camera_pose7, const char* camera_params_json,
                                    const char* class_mapping_json, const char* slot_strings,
                                    const int32_t* slot_string_offsets, int num_slots, int height, int width,
                                    const double* keypoints, const uint8_t* visibility,
                                    const int32_t* person_of_slot, int num_people, int num_joints,
                                    char* out_host, int64_t capacity);

/* f3 / S6: host-side COCO annotations of a batch (no CUDA; HOST pointers).  For every kept record of
 * frames [0, frames) of records_host [B][N] / n_out_host [B] one object
 *   {"id", "image_id", "category_id", "bbox": [x_min, y_min, w, h], "area": count, "iscrowd": 0,
 *    "occlusion", "truncation" (both as Python prints round(x, 6))[, "keypoints": [x, y, v]*J, "num_keypoints"]}
 * joined by ", " — what json.dump(list) writes between its brackets.  image_ids int64 [frames];
 * annotation ids count up from first_annotation_id; keypoints double [B][P][J][2], visibility uint8
 * [B][P][J], person_of_slot int32 [B][N] (-1 = none), all three NULL without skeletons.
 * Returns the number of bytes written or a negative CSPE_ERR_*. */
int64_t cspe_format_coco_host(const cspe_record* records_host, const int32_t* n_out_host, int B, int N, int frames,
                              const int64_t* image_ids, int64_t first_annotation_id, const double* keypoints,
                              const uint8_t* visibility, const int32_t* person_of_slot, int num_people,
                              int num_joints, char* out_host, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* CSPE_H_ */
