// ipa_fold.cu -- second translation unit of ipa.cu: only k_fold_multi, compiled as k_fold_multi_call with the field multiplication
// as an out-of-line call (used for folds with many outputs; see the launch site in ipa.cu for the measurements).
#define HALO_FP_MUL_CALL 1
#define HALO_IPA_FOLD_TU 1
#include "ipa.cu"
