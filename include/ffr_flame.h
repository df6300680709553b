/*
ffr_flame.h -- C ABI of the host-side flame model: flame JSON text -> ffr_flame_desc.

Replaces, for the hot path only, what the reference does at construction time:
  Json(std::istream&)            utils/json.cpp:15-18   (nlohmann parse, // and block comments allowed)
  Flame<dims>::Flame(const Json&) types/flame.hpp:91-210 (validation, _optimize, cumulative weights)
  XForm<dims>::XForm             types/xform.hpp:71-172
  Affine<T,N>::Affine(Json&)     types/affine.hpp:45-92
  Variation::parseVariation      variations/variations.hpp:2387-2626 (+ every constructor's
                                 derived-parameter precompute, evaluated with the host libm
                                 using the same expressions)
The format is FLAME_JSON.md unchanged. Numbers are converted with strtod like nlohmann
does, so every coefficient is the same double as in the reference.
*/

#ifndef FFR_FLAME_H
#define FFR_FLAME_H

#include "ffr_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ffr_flame ffr_flame; /* owns every array the desc points into */

/* Parse + validate + flatten. Returns NULL and a message in err on failure; the
   message follows the reference's JsonError texts where one exists. */
ffr_flame *ffr_flame_from_json(const char *text, size_t len, char *err, size_t errlen);
/* Same with the "size" array overridden (the BASELINE configs re-size the shipped
   examples; the reference has no setter, one edits the JSON). n_size must equal dims. */
ffr_flame *ffr_flame_from_json_sized(const char *text, size_t len, const uint64_t *size,
        int n_size, char *err, size_t errlen);
/* Same for either build of the reference: elem_size 8 = num_t double / hist_t uint64_t (as
   shipped), 4 = float / uint32_t (types/types.hpp:24-41). All constructor arithmetic runs in
   that num_t; size may be NULL. */
ffr_flame *ffr_flame_from_json_ex(const char *text, size_t len, const uint64_t *size,
        int n_size, int elem_size, char *err, size_t errlen);
const ffr_flame_desc *ffr_flame_get_desc(const ffr_flame *f);
void ffr_flame_free(ffr_flame *f);

/* BufferRenderer::_init (buffer_renderer.hpp:114-140): mult_d[i] = size/(hi-lo) *
   (1-2^-52), mult_i = {1,size0,size0*size1}, cells = prod(size). Returns FFR_E_INVALID
   ("histogram too big") when cells >= 2^48. */
int ffr_flame_layout(const ffr_flame_desc *desc, double mult_d[FFR_MAX_DIMS],
        uint64_t mult_i[FFR_MAX_DIMS], uint64_t *cells, uint64_t *cell_size);

/* name <-> opcode of the variation factory (variations.hpp:2387-2612) */
const char *ffr_var_name(uint32_t op);
uint32_t ffr_var_op_from_name(const char *name);

/* The flame as `std::cerr << "flame: " << json_flame` echoes it (ffr_buf.cpp:129 through
   utils/json.cpp:203-207, i.e. nlohmann's compact dump: keys sorted, no whitespace, comments
   gone). Writes at most outlen-1 characters + NUL into out (may be NULL) and returns the full
   length; 0 and a message in err when the text does not parse. */
size_t ffr_flame_json_echo(const char *text, size_t len, char *out, size_t outlen,
        char *err, size_t errlen);

/* ffr-buf's batch-size heuristic (ffr_buf.cpp:94-101): clamp((samples+255)>>8, 4096, 1<<20) */
uint64_t ffr_reference_batch_size(uint64_t samples);

#ifdef __cplusplus
}
#endif

#endif /* FFR_FLAME_H */
