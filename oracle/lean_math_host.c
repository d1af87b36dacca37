/* lean_math_host.c — host instantiation of crnn_b200/csrc/lean_math.h for the oracle's SHARED-MATH mode
 * (crnn_oracle_set_shared_math).  TEST INFRASTRUCTURE.  Built with -ffp-contract=off so that every operation
 * of the header rounds exactly as the CUDA kernels' copy does. */
#include "../crnn_b200/csrc/lean_math.h"

double crnn_lean_log(double x) { return lean_log(x); }
double crnn_lean_exp(double x) { return lean_exp(x); }
double crnn_lean_pow(double x, double y) { return lean_pow(x, y); }
double crnn_lean_log10(double x) { return lean_log10(x); }
double crnn_lean_exp10(double x) { return lean_exp10(x); }
