/* Empty stand-in for <sys/sysctl.h>, which glibc >= 2.32 no longer ships.
 * The reference includes it unconditionally (src/core/parallel.cpp:46) but on
 * Linux only uses sysconf() (parallel.cpp:884-885), so an empty header is
 * behaviour-neutral.  Used only by oracle/Makefile via -Istubs. */
