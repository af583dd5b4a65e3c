/* qaoa_stubs.c — glue for oracle/_ref/libqaoa_ref.so (TEST INFRASTRUCTURE): the reference's algorithms/qaoa.c is compiled
 * unmodified; the only symbol it takes from elsewhere in the reference is its logger. */
void geometric_log_info(const char* fmt, ...) { (void)fmt; }
