/* oracle/spv_ctx.h — TEST INFRASTRUCTURE ONLY.
 * Execution environment of one invocation of a reference SPIR-V module translated by oracle/spv2c.py:
 * descriptor bindings, push constants, stage inputs/outputs and the two services the fixed-function hardware
 * gave the reference's shaders (image sampling, screen-space derivatives). */
#ifndef SPV_CTX_H
#define SPV_CTX_H
#include <stddef.h>
#include <stdint.h>

typedef struct { uint32_t set, binding, index; } spv_handle;        /* an image or sampler descriptor */
typedef struct { spv_handle image, sampler; } spv_sampled;

/* SPV_KHR_ray_query: what OpRayQueryInitializeKHR stored.  The traversal is the environment's (ray-tracing hardware in
 * the reference; the shadow-ray definition of oracle/shadow.c here). */
typedef struct { uint64_t accel; uint32_t flags, cull_mask; float origin[3], t_min, direction[3], t_max; uint32_t initialised; } spv_rayq;

typedef struct spv_ctx {
    struct { uint8_t* ptr; uint64_t size; } buf[4][16];              /* [descriptor set][binding] */
    uint8_t* push;                                                   /* push-constant block */
    void* builtin[64];                                               /* by SPIR-V BuiltIn number (15 FragCoord, 28 GlobalInvocationId ...) */
    void* in_loc[16];                                                /* stage inputs by Location */
    void* out_loc[16];                                               /* stage outputs by Location */
    /* OpImageSample{Implicit,Explicit}Lod: coord has n components; has_lod = 0 means implicit level of detail */
    void (*sample)(struct spv_ctx*, spv_handle image, spv_handle sampler, const float* coord, int n, int has_lod,
                   float lod, float* out4);
    /* OpDPdx / OpDPdy of a value; loc = Location of the stage input it was loaded from, or -1 */
    void (*dpd)(struct spv_ctx*, int is_y, int loc, int n, const float* value, float* out);
    /* OpRayQueryProceedKHR: 1 while a candidate waits for the shader's verdict.  OpRayQueryGetIntersectionTypeKHR on the
     * committed intersection: 0 = none, 1 = triangle. */
    int (*rq_proceed)(struct spv_ctx*, spv_rayq*);
    uint32_t (*rq_committed_type)(struct spv_ctx*, spv_rayq*);
    int killed;                                                      /* OpKill */
    void* user;
} spv_ctx;

typedef void (*spv_entry_fn)(spv_ctx*);
#endif
