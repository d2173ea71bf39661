/*
 * b200nuts_model.h -- device-model plug-in interface of libb200nuts.so.
 *
 * What it replaces.  In the reference a model reaches the sampler through two seams:
 *   - the link-time hook   stan::model::model_base& new_model(stan::io::var_context& data, unsigned int seed,
 *                          std::ostream* msg)                                   (src/cmdstan/command.hpp:71-73)
 *   - the virtual contract stan::model::model_base (stan/src/stan/model/model_base.hpp:36-705), of which the sampling
 *     path uses  model_name :57,  num_params_r (prob_grad.hpp:57),  log_prob_propto_jacobian(var) :191-301 through
 *     stan::model::log_prob_grad / stan::model::gradient (stan/src/stan/model/gradient.hpp:21-35),
 *     write_array :370,  constrained_param_names :149.
 * A C++ virtual cannot be called from a kernel and there is no AD tape on the device, so the device-side equivalent is
 * a table of C functions that ENQUEUE CUDA work: the model's closed-form log density + gradient for a whole batch of
 * chains in one call (the cross-chain formulation is the point of the engine), and write_array for a batch of draws.
 * A model is a shared library built with nvcc that exports
 *
 *     const b200nuts_model_vtable* b200nuts_model_entry(int index);     // index = 0, 1, ...; NULL past the last model
 *
 * and is loaded with b200nuts_load_model_library(), or a vtable registered in-process with b200nuts_register_model().
 * Registered models get ids >= B200NUTS_MODEL_PLUGIN_BASE, usable as b200nuts_config.model; they run in the lock-step
 * engine (every chain is advanced to its next gradient request, then ONE log_prob_grad call serves all of them).
 *
 * Rules a model must keep (they are what makes chains shardable and runs reproducible; the built-in kernels keep them):
 *   - the result for a chain depends only on that chain's q (fixed reduction order; no dependence on the batch);
 *   - all work is enqueued on the given stream; no synchronisation, no allocation in log_prob_grad;
 *   - non-finite lp is allowed (V = +inf  =>  divergent transition, as in the reference).
 */
#ifndef B200NUTS_MODEL_H
#define B200NUTS_MODEL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NUTS_MODEL_ABI 2 /* 2: transform_inits added at the end of the vtable; vtables with abi == 1 are still accepted */
#define B200NUTS_MODEL_PLUGIN_BASE 100

enum b200nuts_dtype { B200NUTS_F64 = 0, B200NUTS_I32 = 1, B200NUTS_I8 = 2 };

/* one named data variable as handed to b200nuts_set_data_*: HOST memory, row-major, valid only during create() */
typedef struct b200nuts_data_view {
  const char* name;
  int32_t dtype; /* enum b200nuts_dtype */
  int32_t nd;
  const int64_t* dims;
  const void* host;
} b200nuts_data_view;

typedef struct b200nuts_model_vtable {
  uint32_t abi;     /* B200NUTS_MODEL_ABI */
  const char* name; /* model_base::model_name() */

  /* new_model(): validate the data (return 65 DATAERR / 66 NOINPUT with a message in err), upload what the kernels need
     to `device`, report the number of unconstrained parameters and of write_array outputs.  *model is passed back to
     every other call. */
  int (*create)(const b200nuts_data_view* data, int n_data, int device, void** model, int64_t* num_params_r,
                int64_t* num_outputs, char* err, int errlen);
  void (*destroy)(void* model);

  /* log_prob_propto_jacobian + gradient for a batch (all pointers are DEVICE pointers):
       n = n_list ? *n_list : n_fixed  requests; request s is chain c = list ? list[s] : s;
       zV[c] = -lp(Q[c*ld .. c*ld + D)),   G[c*ld + i] = -d lp / d q_i.
     jacobian = 1 on the sampling path; 0 only from b200nuts_log_prob_grad(jacobian = 0).
     Returns the number of kernels enqueued (> 0) or a negative value on failure. */
  int (*log_prob_grad)(void* model, const double* Q, double* G, double* zV, const int* list, const int* n_list,
                       int n_fixed, int ld, int jacobian, void* cuda_stream);

  /* write_array(include_tparams, include_gqs without RNG use) for n draws: q[s*ldq .. + D) -> out[s*ldo .. + num_outputs).
     NULL: outputs are the unconstrained parameters themselves (num_outputs must equal num_params_r). */
  int (*write_array)(void* model, const double* q, int64_t n, int64_t ldq, double* out, int64_t ldo, void* cuda_stream);

  /* constrained_param_names(): name of output column k ("beta.3"); NULL: "v<k+1>" */
  int (*output_name)(void* model, int64_t k, char* buf, int buflen);

  /* (abi >= 2; may be NULL) transform_inits (stan/src/stan/model/model_base.hpp:348, :596): the user's initial values on the
     CONSTRAINED scale -> the unconstrained vector, for ONE chain, on the HOST (it runs once per chain at set-up, like the
     reference's serial loop hmc_nuts_diag_e_adapt.hpp:360-362).
       x [num_outputs]: write_array's column order (the parameters are its first columns; transformed parameters and generated
                        quantities that follow are ignored).  NaN = the user gave no value for this element.
       q [num_params_r] out.  For an element the user did not give write b200nuts_init_not_given(code), code = 0 unbounded, 1 lower
                        bound 0, 2 unit interval: the engine then draws it on the unconstrained scale exactly like the reference's
                        random_var_context (initialize.hpp:93-101).  Models with other transforms report "not given" for the
                        whole vector when anything is missing (the engine then uses random inits).
     Returns 0, or 78 (CONFIG) with a message in err when a value violates its constraint (lb_free / lub_free throw there).
     NULL: init=<file> is refused for this model, only init=<radius> works. */
  int (*transform_inits)(void* model, const double* x, double* q, char* err, int errlen);
} b200nuts_model_vtable;

/* the "not given" marker of transform_inits: a quiet NaN whose low payload bits carry the parameter's transform */
static inline double b200nuts_init_not_given(int code) {
  union { uint64_t u; double d; } v;
  v.u = 0x7ff8000000000000ull | (uint64_t)(code & 3);
  return v.d;
}

typedef const b200nuts_model_vtable* (*b200nuts_model_entry_fn)(int index);

/* registers a model; returns its id (>= B200NUTS_MODEL_PLUGIN_BASE) or a negative error code (-64: bad vtable).
   Registering a name twice returns the existing id.  The vtable must outlive every engine that uses it. */
int b200nuts_register_model(const b200nuts_model_vtable* vt);
/* dlopen(path) and register every model of its b200nuts_model_entry; returns the id of the first one, or a negative
   error code (-66: cannot load, -64: no entry point / bad vtable); message via b200nuts_last_error(NULL) */
int b200nuts_load_model_library(const char* path);
/* id of a registered (or built-in) model by name, -1 if unknown */
int b200nuts_find_model(const char* name);
/* constrained_param_names of output column k of an engine's model (built-in or plug-in); valid after the data is set */
struct b200nuts_engine;
int b200nuts_output_name(struct b200nuts_engine* e, int64_t k, char* buf, int buflen);
/* model_base::transform_inits for n chains: x HOST [n][num_outputs] (NaN = not given) -> q HOST [n][num_params_r], ready for
   b200nuts_set_inits.  Built-in models apply their own transforms; plug-in models need vtable.transform_inits (else 64 USAGE).
   A value outside its constraint gives 78 (CONFIG) with the reference's message. */
int b200nuts_transform_inits(struct b200nuts_engine* e, const double* x, int64_t n, double* q);

#ifdef __cplusplus
}
#endif
#endif
