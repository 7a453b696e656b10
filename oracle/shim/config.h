/* oracle/shim/config.h -- TEST INFRASTRUCTURE ONLY.
 * Empty stand-in for the autotools-generated config.h that the reference includes
 * (LP_ompi.h:14).  The build flags that matter (-DHAVE_OPENBLAS) come from oracle/Makefile. */
