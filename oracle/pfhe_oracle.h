/* pfhe_oracle.h -- error codes shared by the oracle (mirror of NttError / RNSError,
 * primus_ntt/src/error.rs:7-49, primus_rns/src/error.rs:7-20). TEST INFRASTRUCTURE ONLY.
 * The function set is declared implicitly by oracle_impl.inc (two instantiations, suffix 32/64)
 * and bound from Python in oracle/oracle.py. */
#ifndef PFHE_ORACLE_H
#define PFHE_ORACLE_H
enum {
    O_OK = 0,
    O_ERR_NO_PRIMITIVE_ROOT = 1,
    O_ERR_DEGREE_CONVERSION = 2,
    O_ERR_DEGREE_TOO_LARGE = 3,
    O_ERR_NTT_TABLE = 4,
    O_ERR_MODULUS_TOO_LARGE = 5,
    O_ERR_RNS_EMPTY = 6,
    O_ERR_RNS_NOT_COPRIME = 7
};
#endif
