/*
 * ref_backend_renamed.c -- TEST INFRASTRUCTURE.  The UNMODIFIED reference
 * backend.c compiled under ref_* names, so that it can live in one process with
 * the new backend (which exports the original names).
 */
#define find_best_match ref_find_best_match
#define set_forward_window ref_set_forward_window
#define get_forward_window ref_get_forward_window
#define set_max_match_count ref_set_max_match_count
#define get_max_match_count ref_get_max_match_count
#define set_magic_factor1 ref_set_magic_factor1
#define get_magic_factor1 ref_get_magic_factor1
#define set_magic_factor2 ref_set_magic_factor2
#define get_magic_factor2 ref_get_magic_factor2
#include "backend.c"
