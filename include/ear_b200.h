/*
 * ear_b200 -- C ABI of the B200-native replacement for E.A.R's ray-tracing render path.
 *
 * The reference (aothms/ear) has no plugin/FFI interface; the seam this library replaces is
 * the C++ call
 *     void Scene::Render(int band, int sound, float absorbtion_factor, int num_samples,
 *                        float dry, const std::vector<Recorder*>& rec, int keyframeID)
 * (src/Scene.h:77, body src/Scene.cpp:111-318) as invoked once per (sound, keyframe, band)
 * "context" by SceneContext::operator() (src/SceneContext.h:46-48) from the boost thread
 * fan-out in Render() (src/EAR.cpp:170-207).  One ear_b200_render() call replaces that whole
 * fan-out.  Results come back the way the reference returns them -- as recorder tracks
 * (FloatBuffer: data, first_sample, real_length; src/Recorder.h:55-92) -- so the host post
 * chain (Power/Truncate/T60/convolution, src/EAR.cpp:210-386) runs unchanged on top.
 *
 * Conventions: plain pointers and sizes, no C++ types, no exceptions across the boundary.
 * Every call returns 0 on success, non-zero on failure; ear_b200_last_error() then holds
 * the text a host rethrows as std::runtime_error so `main` prints the reference's
 * "Error: <what>" line (src/EAR.cpp:404-408).  Host buffers stay owned by the caller;
 * results are owned by the library until ear_b200_result_free().  There is no CPU fallback:
 * without a CUDA device every compute entry point fails.
 *
 * Threading: a scene holds per-scene scratch (ray pool, visibility maps, launch statistics), so calls on ONE scene
 * must not overlap; different scenes (e.g. one per GPU, as the CLI does) may be driven from different host threads.
 * ear_b200_last_error() is per host thread.  Tuning knobs (EAR_B200_* environment variables, README.md) are read
 * when a scene is created.
 *
 * All arithmetic is IEEE float32 with the reference's operation order (no FMA contraction
 * on the geometry path); indices are int32.  Triangle index == position in `verts`, which
 * must be the concatenation of the scene's MESH blocks in file order (src/Scene.cpp:103-106).
 */
#ifndef EAR_B200_H
#define EAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EAR_B200_ABI_VERSION 3        /* 2: ear_b200_context.stream_id, scene images, device post chain
                                         3: mesh-emitter sources, one track per mono recorder in device buffers,
                                            device BVH build, multi-GPU render inside the library */
#define EAR_B200_MAX_BANDS 8          /* the .ear format carries 3; up to 8 through the ABI */
#define EAR_B200_SAMPLE_RATE 44100    /* src/Recorder.h:35 */
#define EAR_B200_MONO 1               /* OUT1, MonoRecorder  (src/MonoRecorder.cpp:83-97)   */
#define EAR_B200_STEREO 2             /* OUT2, StereoRecorder (src/StereoRecorder.cpp:97-130) */
#define EAR_B200_POINT_SOURCE 0       /* AbstractSoundFile without a mesh: Sample_Sphere at its location (src/SoundFile.cpp:222-226) */
#define EAR_B200_MESH_SOURCE 1        /* AbstractSoundFile::mesh set: area-weighted emitter triangle, uniform point, hemisphere
                                         about the triangle's own normal (src/SoundFile.cpp:216-221, src/Mesh.cpp:143-154,
                                         src/Triangle.cpp:44-53); bounce 0 is recorded, no direct lobe (src/Scene.cpp:185,299) */

typedef struct ear_b200_scene ear_b200_scene;    /* triangles + BVH + materials on one GPU */

/* One recorder as seen by one context (positions already evaluated at the context's
 * keyframe: Recorder::getLocation(kf), StereoRecorder::getRightEar(kf)). */
typedef struct ear_b200_recorder {
	int32_t kind;                                /* EAR_B200_MONO | EAR_B200_STEREO */
	float position[3];
	float right_ear[3];                          /* stereo only */
	float head_size;                             /* stereo only */
	float head_absorption[EAR_B200_MAX_BANDS];   /* stereo only; as held after load:
	                                                max(0, (1-a)^4)  (src/StereoRecorder.cpp:55-57) */
} ear_b200_recorder;

/* One SceneContext (src/SceneContext.h:28-45): a (sound, keyframe, band) render. */
typedef struct ear_b200_context {
	int32_t band;                 /* column of the material table */
	int32_t stream_id;            /* 0: the context's position in the call keys its random streams; k > 0: key k - 1 is used
	                                 instead, so a context traced alone or on another GPU follows the paths it has inside
	                                 the full call (the CLI deals contexts to GPUs and passes the global index + 1) */
	int64_t num_samples;          /* rays; the CLI passes settings "samples"/10 (src/EAR.cpp:81) */
	float absorption_factor;      /* 1 - air absorption[band] per metre (src/EAR.cpp:180) */
	float dry_level;              /* direct-sound gain (src/Scene.cpp:308-309) */
	float gain;                   /* source gain; tracks are scaled by gain^2 (src/Scene.cpp:313-314) */
	float source_position[3];     /* AbstractSoundFile::getLocation(kf) (src/SoundFile.cpp:142-148) */
	int32_t source_kind;          /* EAR_B200_POINT_SOURCE | EAR_B200_MESH_SOURCE */
	int32_t emitter_first;        /* mesh source: its triangles are [emitter_first, emitter_first + emitter_count) of the */
	int32_t emitter_count;        /*              table given to ear_b200_scene_set_emitters                             */
	int32_t reserved;
} ear_b200_context;

typedef struct ear_b200_options {
	int32_t max_bounces;          /* reference hard-codes 1000 (src/Scene.cpp:143); 0 -> 1000 */
	int32_t n_bins;               /* bins per track; 0 -> sized from max_bounces x scene diagonal */
	uint64_t seed;                /* Philox4x32-10 key; streams are keyed by (seed, context, ray) */
	int64_t first_ray;            /* shard: this call traces ray ids [first_ray, first_ray+ray_count) */
	int64_t ray_count;            /*        of every context; ray_count < 0 -> all of num_samples    */
	int32_t finalise;             /* 1: apply 1/N, direct sound, gain^2 (src/Scene.cpp:286-316);
	                                 0: leave raw partial sums (multi-GPU: reduce first, then finalise) */
	int32_t reserved;
	float post_exponent;          /* > 0 (with finalise): also run Render()'s post chain on the device before the download --  */
	float post_divisor;           /*   Power(post_exponent), global maximum, Truncate(getLength(maximum / post_divisor)), T60  */
	                              /*   (src/EAR.cpp:209-228; the reference uses 0.335 and 256).  0: tracks come back raw.     */
} ear_b200_options;

/* FloatBuffer view (src/Recorder.h:55-65). `length` is the allocated bin count. */
typedef struct ear_b200_track {
	float* data;
	uint32_t first_sample;
	uint32_t real_length;
	uint32_t length;
	uint32_t reserved;
} ear_b200_track;

typedef struct ear_b200_result {
	int32_t n_contexts;
	int32_t n_recorders;
	ear_b200_track* tracks;       /* [n_contexts][n_recorders][2]; mono uses slot 0 only */
	uint64_t rays;                /* rays traced by this call */
	uint64_t segments;            /* Scene::Bounce invocations (src/Scene.cpp:151), incl. the final miss */
	uint64_t occlusion_queries;   /* Scene::Connect invocations (src/Scene.cpp:194) */
	uint64_t contributions;       /* Recorder::Record invocations (src/Scene.cpp:259) */
	uint64_t bin_updates;         /* track[i] += v operations */
	uint64_t dropped_updates;     /* bin updates beyond n_bins (must be 0 for parity) */
	double device_ms;             /* CUDA-event time of the trace kernels on the library's stream (slowest GPU) */
	double bvh_build_ms;
	float* t60;                   /* post chain only: RecorderTrack::T60 per track, [n_contexts][n_recorders][2]; else NULL */
	float maximum;                /* post chain only: the global maximum after Power() */
	float reserved;
} ear_b200_result;

/* Cumulative per-scene launch statistics of the trace engine (reset by ear_b200_scene_stats_reset).
 * Kernel classes: 0 shade/refill/enqueue, 1 closest-hit traversal, 2 BVH any-hit traversal, 3 splat,
 * 4 list binning (scan, scatter), 5 finalise (scale, direct), 6 visibility-map lookups, 7 visibility-map builds.
 * ms[] are CUDA-event times on the launch stream. */
typedef struct ear_b200_stats {
	uint64_t launches[8];
	double ms[8];
	uint64_t iterations;
	uint64_t reserved;
} ear_b200_stats;

const char* ear_b200_last_error(void);
int32_t ear_b200_abi_version(void);
int32_t ear_b200_device_count(void);
/* The library keeps device memory it has released (ray pools, BVH scratch, visibility maps: GBs) in a per-process cache so
 * that the next scene or render does not pay cudaMalloc / cudaFree again (EAR_B200_CACHE_MB caps it, default 32768).
 * This returns all of it to the driver. */
void ear_b200_release_cached_memory(void);

/* Uploads the triangle soup (file order), builds the BVH, keeps everything resident on `device`.
 * materials: [n_materials][n_bands][4] = {reflection, transmission, kept, specularity} where
 * kept = absorption_coefficient as derived by src/Material.cpp:33-60.  Replaces Scene::addMesh /
 * Mesh::Combine / Triangle ctor (src/Scene.cpp:103-106, src/Mesh.cpp:116-123, src/Triangle.cpp:26-43). */
int32_t ear_b200_scene_create(const float* verts /*[n_tris][3][3]*/, const int32_t* tri_material /*[n_tris]*/,
                              int32_t n_tris, const float* materials, int32_t n_materials, int32_t n_bands,
                              int32_t device, ear_b200_scene** out);
void ear_b200_scene_destroy(ear_b200_scene* scene);
/* Emitter triangles of the scene's mesh sources (the `mesh` block inside a sound source, src/SoundFile.cpp:50-53):
 * all of them concatenated, file order; contexts of mesh sources name a range of this table.  They emit only (they
 * are not part of the geometry rays hit).  Normals and areas are formed as Triangle's constructor does
 * (src/Triangle.cpp:26-34).  Replaces the table of an earlier call; n = 0 clears it. */
int32_t ear_b200_scene_set_emitters(ear_b200_scene* scene, const float* verts /*[n][3][3]*/, int32_t n);

/* Multi-GPU replication: the acceleration data of a scene is one contiguous device
 * image.  One rank builds the scene, writes its image into a device buffer it can broadcast (NCCL over
 * NVLink), and every other rank adopts the received bytes -- one host BVH build per job instead of one per
 * GPU.  The reference has nothing to replace here (all its threads share one Scene, src/Scene.cpp:283-312);
 * this is the cross-process form of that sharing.
 *   image_size   bytes the image occupies
 *   image_write  copies the image to `dst_device` (device memory on any GPU of this process, >= image_size)
 *   create_from_image  new scene on `device` from image bytes resident in device memory; the bytes are copied,
 *                      the source buffer may be freed afterwards.  Fails on a foreign or truncated image. */
int32_t ear_b200_scene_image_size(ear_b200_scene* scene, uint64_t* bytes);
int32_t ear_b200_scene_image_write(ear_b200_scene* scene, void* dst_device, uint64_t bytes);
int32_t ear_b200_scene_create_from_image(const void* src_device, uint64_t bytes, int32_t device, ear_b200_scene** out);
/* Same process, several GPUs: a copy of `scene` on `device` (peer copy of the image, no second BVH build). */
int32_t ear_b200_scene_clone(ear_b200_scene* scene, int32_t device, ear_b200_scene** out);

/* Harness: Mesh::RayIntersection (src/Mesh.cpp:33-56) for explicit rays.
 * tri_index[i] = winning triangle or -1; t[i] = its distance (undefined on miss). */
int32_t ear_b200_first_hit(ear_b200_scene* scene, const float* origins /*[n][3]*/, const float* dirs /*[n][3]*/,
                           int64_t n, int32_t* tri_index, float* t);
/* Harness: Scene::Connect / Mesh::LineIntersection (src/Scene.cpp:84-96, src/Mesh.cpp:58-71)
 * for explicit segments p -> x.  out[i] = 1 if occluded. */
int32_t ear_b200_occluded(ear_b200_scene* scene, const float* p /*[n][3]*/, const float* x /*[n][3]*/,
                          int64_t n, uint8_t* out);
/* Harness: replays ray ids [first_ray, first_ray+n) of one context through the bounce loop and
 * returns, per ray, the triangle hit at each bounce (-1 terminates) -- used to check emission,
 * Material::Bounce and Sample_Hemi against the oracle path by path.  hits: [n][max_bounces]. */
int32_t ear_b200_trace_paths(ear_b200_scene* scene, const ear_b200_context* ctx, int32_t ctx_index,
                             const ear_b200_options* opt, int64_t n, int32_t* hits, float* final_state /*[n][8]*/);

/* The hot path.  Host buffers in, host tracks out (device copies inside).
 * rec: [n_contexts][n_recorders].  The tracks of one result live in a single page-locked host block owned by the
 * library (a released block is kept for the next result of about that size; ear_b200_release_cached_memory()
 * returns it to the driver); they stay valid, and may be written, until ear_b200_result_free(). */
int32_t ear_b200_render(ear_b200_scene* scene, const ear_b200_context* ctx, int32_t n_contexts,
                        const ear_b200_recorder* rec, int32_t n_recorders, const ear_b200_options* opt,
                        ear_b200_result** out);
void ear_b200_result_free(ear_b200_result* result);

/* Several GPUs in ONE process (the CLI's form).  A group is `scene` plus peer copies of its device
 * image on the other listed GPUs (devices[0] must be the scene's own device; the scene stays the caller's, the copies
 * are destroyed with the group).  ear_b200_group_render is ear_b200_render over all of them: every GPU traces a
 * contiguous share of the ray ids of EVERY context into its own partial histogram (one host thread per GPU); the
 * partials meet on GPU 0 in one sliced reduce -- GPU g sums slice g of all partials through peer loads over
 * NVLink / NVSwitch and stores it into GPU 0's buffer -- then GPU 0 finalises (and runs the post chain if asked) and
 * the tracks are downloaded.  Random streams are keyed by (seed, context, ray id): the result does not depend on the
 * number of GPUs beyond the order of the float sums. */
typedef struct ear_b200_group ear_b200_group;
int32_t ear_b200_group_create(ear_b200_scene* scene, const int32_t* devices, int32_t n_devices, ear_b200_group** out);
void ear_b200_group_destroy(ear_b200_group* group);
int32_t ear_b200_group_size(ear_b200_group* group);
int32_t ear_b200_group_render(ear_b200_group* group, const ear_b200_context* ctx, int32_t n_contexts,
                              const ear_b200_recorder* rec, int32_t n_recorders, const ear_b200_options* opt,
                              ear_b200_result** out);

/* Device-resident variant for one-process-per-GPU sharding: accumulates raw partial sums into
 * caller-owned device memory so the caller can reduce them across ranks (NCCL) before finalising.
 *   d_hist   float  [n_contexts][n_recorders][tpr][n_bins]   (zeroed by the caller)
 *   d_range  uint32 [n_contexts][n_recorders][tpr][2]        {first_sample (init 132299), real_length (init 0)}
 *            tpr = ear_b200_tracks_per_recorder(rec, n_contexts * n_recorders): 2 when any recorder of the call is
 *            stereo, else 1 (64 mono recorders do not carry 64 dead tracks through memset, atomics and the reduce)
 *   d_counters uint64 [8]: rays, segments, occlusion_queries, contributions, bin_updates, dropped, -, -
 * `stream` is a cudaStream_t (0 = legacy default stream).  All work is ordered on `stream`; the call returns once
 * the last ray has ended (the wavefront loop reads a 32-byte counter block every few iterations to know when to
 * stop), so the results are complete -- but not yet visible to other streams -- on return. */
int32_t ear_b200_trace_device(ear_b200_scene* scene, const ear_b200_context* ctx, int32_t n_contexts,
                              const ear_b200_recorder* rec, int32_t n_recorders, const ear_b200_options* opt,
                              int32_t n_bins, float* d_hist, uint32_t* d_range, uint64_t* d_counters, void* stream);
/* K7 on device buffers: x 1/N, direct sound, x gain^2 (src/Scene.cpp:286-316). Asynchronous. */
int32_t ear_b200_finalise_device(ear_b200_scene* scene, const ear_b200_context* ctx, int32_t n_contexts,
                                 const ear_b200_recorder* rec, int32_t n_recorders, int32_t n_bins,
                                 float* d_hist, uint32_t* d_range, void* stream);
/* Tracks each recorder owns in the device buffers of a call with these recorders (see ear_b200_trace_device). */
int32_t ear_b200_tracks_per_recorder(const ear_b200_recorder* rec, int32_t n);
/* Bins per track the library would choose for these options (same rule as ear_b200_render). */
int32_t ear_b200_default_bins(ear_b200_scene* scene, const ear_b200_options* opt);
/* SURVEY.md 8(f) rank 1 -- impulse-response convolution, RecorderTrack::Process (src/Recorder.cpp:247-292,
 * direct form, the reference's default build).  out[k] for k in [0, out_len) receives, bit for bit, what the
 * reference's loops leave in the result track: for i in [0, n_dry) (outer) and j in [first, len) (inner)
 *     out[i + offset + j] += dry[i] * p(j)
 * with p(j) = response[j]                                  when response2 == NULL  (static scene, :247-263), or
 *      p(j) = (1 - i/n_dry) * response[j] + (i/n_dry) * response2[j]   keyframe cross-fade (:267-292),
 *      first = min(first_sample, first_sample2), len = max(real_length, real_length2), tracks read as 0 beyond
 *      their allocated length.  Every output sample is accumulated in increasing i, one float multiply and one
 *      float add per term (no FMA), exactly the order the CPU loop produces.
 * Host buffers in and out; out_first / out_real receive the result track's first_sample / real_length. */
int32_t ear_b200_convolve(int32_t device, const float* response, uint32_t length, uint32_t first_sample,
                          uint32_t real_length, const float* response2, uint32_t length2, uint32_t first_sample2,
                          uint32_t real_length2, const float* dry, uint32_t n_dry, uint32_t offset, float* out,
                          uint32_t out_len, uint32_t* out_first, uint32_t* out_real);

/* The same convolution in the frequency domain -- the reference's optional USE_FFTW build of RecorderTrack::Process
 * (src/Recorder.cpp:145-243): zero-padded real FFTs (cuFFT), one complex multiply, one inverse transform.  For keyframed
 * scenes the dry signal is faded out / in before the two transforms, which is the same linear map as the direct form's
 * per-sample interpolation of the two responses.  Same arguments and bookkeeping as ear_b200_convolve; the samples agree
 * with the direct form to float32 FFT accuracy (tests/test_convolve_gpu.py states the tolerance: 2e-5 of the peak), not bit
 * for bit -- the CLI uses it only when EAR_CONVOLUTION=fft is set. */
int32_t ear_b200_convolve_fft(int32_t device, const float* response, uint32_t length, uint32_t first_sample,
                              uint32_t real_length, const float* response2, uint32_t length2, uint32_t first_sample2,
                              uint32_t real_length2, const float* dry, uint32_t n_dry, uint32_t offset, float* out,
                              uint32_t out_len, uint32_t* out_first, uint32_t* out_real);

/* SURVEY.md section 8(f) rank 2 -- the post chain of Render() (src/EAR.cpp:209-228) on device-resident tracks, in
 * the two phases the reference runs it in (every track is compressed before the global maximum is known):
 *   post_power     Recorder::Power(exponent) in place on every track (FloatBuffer::Power, src/Recorder.cpp:101-106:
 *                  sign(x) * |x|^exponent over [first_sample, real_length)), then FloatBuffer::Maximum of each track
 *                  (:76-83).  maximum = the largest of them (host, may be NULL); track_maximum [n_contexts][n_recorders][tpr]
 *                  (host, may be NULL).  Multi-GPU callers reduce `maximum` by MAX before phase 2.
 *   post_truncate  Recorder::Truncate(Recorder::getLength(threshold)) for every recorder (src/Recorder.cpp:399-430,
 *                  108-118; the reference passes threshold = maximum / 256) -- updates real_length in d_range -- and
 *                  RecorderTrack::T60 (src/Recorder.cpp:303-340) of every truncated track into t60 [n_contexts][n_recorders][tpr]
 *                  (host; `EAR calc T60` prints t60[0]).
 * d_hist / d_range as in ear_b200_trace_device, after ear_b200_finalise_device.  Both calls synchronise `stream`. */
int32_t ear_b200_post_power_device(ear_b200_scene* scene, const ear_b200_recorder* rec, int32_t n_contexts, int32_t n_recorders,
                                   int32_t n_bins, float* d_hist, const uint32_t* d_range, float exponent, float* maximum,
                                   float* track_maximum, void* stream);
int32_t ear_b200_post_truncate_device(ear_b200_scene* scene, const ear_b200_recorder* rec, int32_t n_contexts, int32_t n_recorders,
                                      int32_t n_bins, const float* d_hist, uint32_t* d_range, float threshold, float* t60,
                                      void* stream);

/* Launch counts and device time per kernel class since the last reset (synchronises the scene's last stream). */
int32_t ear_b200_scene_stats(ear_b200_scene* scene, ear_b200_stats* out);
void ear_b200_scene_stats_reset(ear_b200_scene* scene);

#ifdef __cplusplus
}
#endif
#endif
