/* refb_stubs.c — glue for oracle/_ref/libqgt_refb.so (TEST INFRASTRUCTURE): the reference's
 * core/quantum_circuit_operations.c is compiled unmodified; these are the four symbols it references from other
 * reference files, none of them on the create / build / execute / destroy path the parity tests drive (they are reached
 * only from the phase-estimation and annealing helpers and from circuits carrying a geometric state or graph). */
#include <stdio.h>
#include <stdlib.h>
static void unreachable(const char* what) { fprintf(stderr, "libqgt_refb: %s is outside the oracle's path\n", what); abort(); }
void* quantum_system_create(size_t a, int b) { (void)a; (void)b; unreachable("quantum_system_create"); return NULL; }
void quantum_system_destroy(void* s) { (void)s; unreachable("quantum_system_destroy"); }
void destroy_computational_graph(void* g) { (void)g; unreachable("destroy_computational_graph"); }
void geometric_destroy_state(void* s) { (void)s; unreachable("geometric_destroy_state"); }
