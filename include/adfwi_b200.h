/*
 * adfwi_b200.h -- C ABI of libadfwi_b200.so: sm_100a CUDA implementation of the wave-propagation
 * hot path of liufeng2317/ADFWI (ADFWI/propagator/{acoustic,elastic}_kernels.py).
 *
 * The reference has no FFI of its own: its hot path is the pair of Python free functions
 * `forward_kernel` (acoustic_kernels.py:179, elastic_kernels.py:917) whose body is the TorchScript
 * time loop `step_forward*` plus the autograd tape PyTorch records through it.  Each entry point
 * below replaces one such (implicit) unit; the citation says which.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - all arrays are dense row-major fp32 (indices int64), [z][x] with x fastest;
 *   - the caller owns every buffer including the workspace; the library never allocates or
 *     frees device memory; all work is enqueued on `stream` (a cudaStream_t passed as void*), no
 *     implicit synchronisation.  Process-wide state is limited to diagnostics and caches that do not
 *     change results: the launch counter, the optional sampled kernel timing (adfwi_timing_*),
 *     per-device caches of the SM count / kernel attributes, and the A/B environment switches
 *     ADFWI_B200_{PDL,EL_LEAN,EL_SPLIT,EL_ADJ_SPLIT} which are read once per process;
 *   - return value: 0 = ok, <0 = ADFWI_E_* (argument error, detected before any launch),
 *     >0 = cudaError_t of a failed launch.  adfwi_strerror() maps either to text.
 */
#ifndef ADFWI_B200_H
#define ADFWI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADFWI_ABI_VERSION 1

enum {
    ADFWI_OK = 0,
    ADFWI_E_NULL = -1,       /* required pointer is NULL */
    ADFWI_E_DIMS = -2,       /* grid / counts out of the supported range */
    ADFWI_E_WORKSPACE = -3,  /* workspace smaller than adfwi_*_workspace_bytes() */
    ADFWI_E_ORDER = -4,      /* fd order not in {4,6} */
    ADFWI_E_MODE = -5        /* backward called on a forward-only descriptor, etc. */
};

/* ------------------------------------------------------------------------------------------
 * Iso-acoustic (p,u,w) solver.  Replaces acoustic_kernels.py:41-176 (step_forward, the time
 * loop) and the autograd tape through it (acoustic_kernels.py:268-278 + loss.backward()).
 * The coefficient planes are those of acoustic_kernels.py:257-265, computed by the caller.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t nzp, nxp;        /* padded grid: nz+2*nabc, nx+2*nabc   (acoustic_kernels.py:230-231) */
    int32_t ns, nt, nr;      /* shots in this call, time steps, receivers */
    int32_t nabc;            /* absorbing-layer width; free_surface_start = nabc (FS) or 1 (:228) */
    int32_t free_surface;    /* 0/1 */
    float   dt;              /* f32(dt): source term is f32(dt)*src_v (:131) */
    float   c1, c2;          /* f32(9/8), f32(-1/24) (:253-254) */
    int32_t n_segments;      /* reference's checkpoint_segments: only shapes the literal
                                semantics of forward_wavefield_u/w (:172-174, :286-288) */
    int32_t save_history;    /* 0 = forward modelling only; 1 = keep what backward needs */
    int32_t ckpt_interval;   /* K: state checkpoint every K steps, stencil history kept for K
                                steps at a time (recomputed in backward).  K<=0 or K>=nt =
                                store-all (no recomputation). */
    int32_t need_g_alpha2;   /* 1 if d/d(alpha2) is wanted (density gradient) */
    int32_t shots_per_group; /* shots advanced per kernel launch; 0 = library picks (all) */
    int32_t reserved[4];     /* [0] bit 0: force the generic (unfused) kernels, bit 1: never use the cluster-persistent
                                small-grid kernels; [1]: shots one CTA of the fused kernels walks through per tile
                                (0 = library picks); [2]: the same for the adjoint kernel only; rest 0 */
} adfwi_acoustic_desc;

/* bytes of workspace forward(+backward) needs for this descriptor (0 on invalid desc) */
size_t adfwi_acoustic_workspace_bytes(const adfwi_acoustic_desc* desc);
/* shots one kernel launch advances for this descriptor (the library's L2-residency grouping) */
int adfwi_acoustic_group_size(const adfwi_acoustic_desc* desc);

/*
 * Forward sweep = acoustic_kernels.py:113-174 for nt steps from a zero state.
 *   alpha1,alpha2,kappa1,kappa2,kappa3 : [nzp][nxp]
 *   src_v [ns][nt]; src_x,src_z [ns]; rcv_x,rcv_z [nr]  -- PADDED grid indices (already +nabc)
 *   rcv_p,rcv_u,rcv_w [ns][nt][nr]  (out; rcv_u / rcv_w may be NULL to skip)
 *   illum_p,illum_u,illum_w [nzp-2nabc][nxp-2nabc] (out, each nullable):
 *       the reference's forward_wavefield_{p,u,w} (literal semantics)
 */
int adfwi_acoustic_forward(const adfwi_acoustic_desc* desc,
                           const float* alpha1, const float* alpha2, const float* kappa1,
                           const float* kappa2, const float* kappa3,
                           const float* src_v, const int64_t* src_x, const int64_t* src_z,
                           const int64_t* rcv_x, const int64_t* rcv_z,
                           float* rcv_p, float* rcv_u, float* rcv_w,
                           float* illum_p, float* illum_u, float* illum_w,
                           void* workspace, size_t workspace_bytes, void* stream);

/*
 * Adjoint sweep = what autograd computes for the tape of the same loop: given the record
 * cotangents g_rcv_* [ns][nt][nr] (each nullable = zero) accumulate
 *   g_alpha1, g_alpha2 [nzp][nxp]  (OVERWRITTEN; summed over the ns shots; g_alpha2 nullable)
 *   g_src_v [ns][nt]               (nullable)
 * `workspace` must be the buffer the matching forward call (save_history=1) filled.
 */
int adfwi_acoustic_backward(const adfwi_acoustic_desc* desc,
                            const float* alpha1, const float* alpha2, const float* kappa1,
                            const float* kappa2, const float* kappa3,
                            const float* src_v, const int64_t* src_x, const int64_t* src_z,
                            const int64_t* rcv_x, const int64_t* rcv_z,
                            const float* g_rcv_p, const float* g_rcv_u, const float* g_rcv_w,
                            float* g_alpha1, float* g_alpha2, float* g_src_v,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * 2-D P-SV elastic solver (iso / VTI / HTI through Cij planes).  Replaces
 * elastic_kernels.py:223-423 / :426-579 (step_forward_PML_{4,6}order),
 * :586-777 / :781-912 (step_forward_ABL_{4,6}order) and the autograd tape through them.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t nzp, nxp;        /* padded grid incl. the fs_offset rows (elastic_kernels.py:286-287) */
    int32_t ns, nt, nr;
    int32_t nz, nx, nabc;    /* physical grid + layer width (illumination crop, :414-418) */
    int32_t free_surface;    /* 0/1 */
    int32_t fd_order;        /* 4 or 6 (anything else is rejected; the Python shim maps !=4 -> 6) */
    int32_t abc_pml;         /* 1 = split-field PML (:223), 0 = multiplicative sponge/ABL (:586) */
    float   dt, dx, dz;      /* f32 roundings of the python floats */
    float   dt_dx, dt_dz;    /* f32(dt/dx), f32(dt/dz) (division done in double, :325-326) */
    float   half_dt;         /* f32(0.5*dt) (:313-316) */
    float   fdc[3];          /* DiffCoef(NN,'s') bits (:20-58) */
    int32_t n_segments;      /* n_segments: shapes the illumination maps (:1017-1021) */
    int32_t save_history;
    int32_t ckpt_interval;
    int32_t shots_per_group;
    int32_t reserved[4];     /* [0] bit 0: force the generic (unfused) kernels; [1]: shots one CTA of the fused kernels walks
                                through per tile (0 = library picks); [2]: the same for the reverse kernels only; rest 0 */
} adfwi_elastic_desc;

size_t adfwi_elastic_workspace_bytes(const adfwi_elastic_desc* desc);

/*
 * coef  : 6 planes [nzp][nxp] in the order C11,C13,C33,C55,bx,bz (C15 = C35 = 0 for every model
 *         the reference can build, parameters.py:38-44; adding 0*x is exact, so they are dropped)
 * bcx,bcz [nzp][nxp] (PML) or damp [nzp][nxp] in bcx with bcz=NULL (ABL)
 * mt    : [ns][3][3] moment tensors;  src_v [ns][nt];  indices are PADDED grid indices
 * rcv   : 5 record arrays [ns][nt][nr] in the order txx,tzz,txz,vx,vz (out)
 * illum : 5 maps [nz][nx], same order (out, nullable as a whole)
 */
int adfwi_elastic_forward(const adfwi_elastic_desc* desc, const float* const* coef,
                          const float* bcx, const float* bcz, const float* mt,
                          const float* src_v, const int64_t* src_x, const int64_t* src_z,
                          const int64_t* rcv_x, const int64_t* rcv_z,
                          float* const* rcv, float* const* illum,
                          void* workspace, size_t workspace_bytes, void* stream);

/* g_rcv: 5 cotangent arrays (entries nullable); g_coef: 6 planes [nzp][nxp] (OVERWRITTEN);
 * g_src_v nullable. */
int adfwi_elastic_backward(const adfwi_elastic_desc* desc, const float* const* coef,
                           const float* bcx, const float* bcz, const float* mt,
                           const float* src_v, const int64_t* src_x, const int64_t* src_z,
                           const int64_t* rcv_x, const int64_t* rcv_z,
                           const float* const* g_rcv, float* const* g_coef, float* g_src_v,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Gradient post-processing.  Replaces GradProcessor.forward
 * (ADFWI/propagator/gradient_process.py:88-135: mute taper :97-98 / grad_taper :51-72, mask :101-107,
 * illumination preconditioner :117-122 with smooth2d :30-49, smoothing :125-131, max-normalisation
 * :134-135) and the device->host->device round trip around it (ADFWI/fwi/acoustic_fwi.py:171-177).
 * The reference computes in numpy float64 on the host; this entry point does the same arithmetic in
 * float64 on the device (the Gaussian filter of smooth2d is applied as its two 1-D factors).
 * `grad` is the float32 gradient plane [nz][nx] (what `model.vp.grad` holds), `forw` the float32
 * illumination plane or NULL, `mask` a float64 plane or NULL.  `out` receives nz*nx doubles.
 * `*out_is_f32_host` (host pointer, may be NULL) is set to 1 when numpy's promotion rules make the
 * reference return float32 (no illumination division and no whole-plane smoothing): the values in
 * `out` are then exactly representable in float32.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t nz, nx;
    int32_t grad_mute;        /* rows / columns of the taper, 0 = none (:97) */
    int32_t grad_smooth;      /* span of the final smoothing, 0 = none (:125) */
    int32_t taper_marine;     /* 1: marine_or_land in {'Marine','Offshore'} (grad_taper's own, case-sensitive test :55) */
    int32_t smooth_below_mute;/* 1: marine_or_land in {'marine','offshore'}: smooth rows >= grad_mute only (:127-128) */
    int32_t norm_grad;        /* :134 */
    int32_t use_illumination; /* forw_illumination and forw given (:117) */
    int32_t illum_span;       /* 40, or min(nz,nx)/2 on small grids (:112-115) */
    int32_t reserved;
    double  thred;            /* 0.0 marine/offshore, 0.001 land/onshore (:90-95) */
    double  vmax;             /* value of the (float32) scalar the reference passes */
} adfwi_gradproc_desc;

size_t adfwi_gradproc_workspace_bytes(const adfwi_gradproc_desc* desc);
int adfwi_gradproc_forward(const adfwi_gradproc_desc* desc, const float* grad, const float* forw, const double* mask,
                           double* out, int* out_is_f32_host, void* workspace, size_t workspace_bytes, void* stream);
/* smooth2d alone (gradient_process.py:30-49) on a float64 plane; workspace as above with nz, nx set */
int adfwi_gradproc_smooth2d(int nz, int nx, int span, const double* in, double* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Record post-processing.  Replaces, for one shot batch, the per-trace max-abs
 * normalisation of the synthetic records (ADFWI/fwi/acoustic_fwi.py:149-150: syn / max_t |syn|, keepdim over
 * time; elastic_fwi.py:224-255) followed by the misfit and by what autograd derives from both:
 *   kind 0  Misfit_waveform_L2.forward         (ADFWI/fwi/misfit/L2.py:22-28):   sum_traces sqrt(sum_t (obs - syn)^2 dt)
 *   kind 1  Misfit_global_correlation.forward  (ADFWI/fwi/misfit/GlobalCorrelation.py:43-70)
 * syn, obs: [ns][nt][nr] float32 (obs already normalised by the caller, as the reference does once at
 * construction, acoustic_fwi.py:68-70).  forward() writes the scalar loss (device float) and leaves the
 * per-trace coefficients of the adjoint source in the workspace; adjoint_source() writes
 * g_syn = grad_loss * d(loss)/d(syn) [ns][nt][nr] (grad_loss: device float, NULL = 1).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t ns, nt, nr;
    int32_t kind;             /* 0 = L2 waveform, 1 = global correlation */
    int32_t normalize;        /* 1 = per-trace max-abs normalisation of syn first */
    int32_t reserved[3];
    double  dt;               /* the misfit's dt (the examples pass 1 or the sampling interval) */
} adfwi_misfit_desc;

size_t adfwi_misfit_workspace_bytes(const adfwi_misfit_desc* desc);
int adfwi_misfit_forward(const adfwi_misfit_desc* desc, const float* syn, const float* obs, float* loss,
                         void* workspace, size_t workspace_bytes, void* stream);
int adfwi_misfit_adjoint_source(const adfwi_misfit_desc* desc, const float* syn, const float* obs, const float* grad_loss,
                                float* g_syn, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Model regularisers.  Replaces Regularization.forward of
 *   kind 0  TV_1order        (ADFWI/fwi/regularization/tv_1order.py:25-52)
 *   kind 1  Tikhonov_1order  (tikhonov_1order.py:22-52)
 *   kind 2  TV_2order        (tv_2order.py:25-52)
 *   kind 3  Tikhonov_2order  (tikhonov_2order.py:26-55)
 * and its autograd derivative.  m: [nz][nx] float32; dx, dz in metres (the reference converts to km);
 * alphax, alphaz: the factors AFTER the caller's step decay (regular_StepLR, base.py:16-18).
 * forward() writes the scalar value (device float) and keeps what backward() needs in the workspace.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t nz, nx;
    int32_t kind;
    int32_t reserved;
    double  dx, dz, alphax, alphaz;
} adfwi_regularization_desc;

size_t adfwi_regularization_workspace_bytes(const adfwi_regularization_desc* desc);
int adfwi_regularization_forward(const adfwi_regularization_desc* desc, const float* m, float* value,
                                 void* workspace, size_t workspace_bytes, void* stream);
int adfwi_regularization_backward(const adfwi_regularization_desc* desc, const float* m, const float* grad_value, float* g_m,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Model-side producers of the elastic coefficient planes.
 *
 * adfwi_elastic_moduli_*: replaces thomsen_to_elastic_moduli (ADFWI/model/parameters.py:71-107), the TI fill-in
 * C55 = C44 / HTI swap (:156-181), b = 1/rho (:47-69) and parameter_staggered_grid (:184-213) -- and their autograd
 * mirror.  vp, vs, rho, eps, delta: [nz][nx].  planes: C11, C13, C33 [nz][nx]; C55 [nz-2][nx-2]; bx [nz][nx-1];
 * bz [nz-1][nx] (the ragged shapes the reference hands to forward_kernel).  backward: g_planes in the same shapes,
 * outputs [nz][nx], each nullable.  gamma does not reach the P-SV planes (C66 is unused there).
 *
 * adfwi_elastic_pad_*: replaces the six pad_torchSingle calls of forward_kernel
 * (ADFWI/propagator/elastic_kernels.py:176-216, :935-946) plus the zero extension to the full grid that the region
 * slices of :303-310 imply: planes in their own shapes -> six full [nzp][nxp] planes; backward = transpose.
 * pml = nabc, top = fs_offset (free surface) or fs_offset + nabc.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int32_t nz, nx; int32_t hti; int32_t reserved; } adfwi_elastic_moduli_desc;
typedef struct { int32_t nz, nx, nzp, nxp, pml, top; int32_t reserved[2]; } adfwi_elastic_pad_desc;

int adfwi_elastic_moduli_forward(const adfwi_elastic_moduli_desc* desc, const float* vp, const float* vs, const float* rho,
                                 const float* eps, const float* delta, float* const* planes, void* stream);
int adfwi_elastic_moduli_backward(const adfwi_elastic_moduli_desc* desc, const float* vp, const float* vs, const float* rho,
                                  const float* eps, const float* delta, const float* const* g_planes,
                                  float* g_vp, float* g_vs, float* g_rho, float* g_eps, float* g_delta, void* stream);
int adfwi_elastic_pad_forward(const adfwi_elastic_pad_desc* desc, const float* const* planes, float* const* full, void* stream);
int adfwi_elastic_pad_backward(const adfwi_elastic_pad_desc* desc, const float* const* g_full, float* const* g_planes, void* stream);

/* misc */
const char* adfwi_strerror(int code);
int adfwi_abi_version(void);
/* number of kernels the library has launched in this process (bench.py's gpu_launches) */
uint64_t adfwi_launch_count(void);


/* Diagnostics for the bench roofline: when every_n > 0, every n-th launch of each kernel class is
 * bracketed by CUDA events on its stream (process-wide switch, off by default, every_n = 0 turns
 * it off).  adfwi_timing_collect() synchronises the recorded events, writes per-class summed
 * milliseconds and sample counts (arrays of length n), clears the samples and returns the number
 * of kernel classes.  Class ids: see enum KernelClass in csrc/common.cuh / adfwi_b200/_lib.py. */
void adfwi_timing_enable(int every_n);
int adfwi_timing_collect(float* ms_sum_host, int* count_host, int n);

#ifdef __cplusplus
}
#endif
#endif /* ADFWI_B200_H */
