// Generated comparator lists: Batcher odd-even merge sort for 16 keys (63 compare-exchanges,
// verified with the 0-1 principle) and the 32-exchange bitonic merge that re-sorts max(A[i], B[15-i]).
// SAP_CE(a, b) must leave max in a and min in b, so the arrays end up in DESCENDING order.
#pragma once
#define SAP_SORT16(v) \
SAP_CE(v[0], v[1]); SAP_CE(v[2], v[3]); SAP_CE(v[0], v[2]); SAP_CE(v[1], v[3]); SAP_CE(v[1], v[2]); SAP_CE(v[4], v[5]); \
SAP_CE(v[6], v[7]); SAP_CE(v[4], v[6]); SAP_CE(v[5], v[7]); SAP_CE(v[5], v[6]); SAP_CE(v[0], v[4]); SAP_CE(v[2], v[6]); \
SAP_CE(v[2], v[4]); SAP_CE(v[1], v[5]); SAP_CE(v[3], v[7]); SAP_CE(v[3], v[5]); SAP_CE(v[1], v[2]); SAP_CE(v[3], v[4]); \
SAP_CE(v[5], v[6]); SAP_CE(v[8], v[9]); SAP_CE(v[10], v[11]); SAP_CE(v[8], v[10]); SAP_CE(v[9], v[11]); SAP_CE(v[9], v[10]); \
SAP_CE(v[12], v[13]); SAP_CE(v[14], v[15]); SAP_CE(v[12], v[14]); SAP_CE(v[13], v[15]); SAP_CE(v[13], v[14]); SAP_CE(v[8], v[12]); \
SAP_CE(v[10], v[14]); SAP_CE(v[10], v[12]); SAP_CE(v[9], v[13]); SAP_CE(v[11], v[15]); SAP_CE(v[11], v[13]); SAP_CE(v[9], v[10]); \
SAP_CE(v[11], v[12]); SAP_CE(v[13], v[14]); SAP_CE(v[0], v[8]); SAP_CE(v[4], v[12]); SAP_CE(v[4], v[8]); SAP_CE(v[2], v[10]); \
SAP_CE(v[6], v[14]); SAP_CE(v[6], v[10]); SAP_CE(v[2], v[4]); SAP_CE(v[6], v[8]); SAP_CE(v[10], v[12]); SAP_CE(v[1], v[9]); \
SAP_CE(v[5], v[13]); SAP_CE(v[5], v[9]); SAP_CE(v[3], v[11]); SAP_CE(v[7], v[15]); SAP_CE(v[7], v[11]); SAP_CE(v[3], v[5]); \
SAP_CE(v[7], v[9]); SAP_CE(v[11], v[13]); SAP_CE(v[1], v[2]); SAP_CE(v[3], v[4]); SAP_CE(v[5], v[6]); SAP_CE(v[7], v[8]); \
SAP_CE(v[9], v[10]); SAP_CE(v[11], v[12]); SAP_CE(v[13], v[14]);
#define SAP_BITONIC_MERGE16(v) \
SAP_CE(v[0], v[8]); SAP_CE(v[1], v[9]); SAP_CE(v[2], v[10]); SAP_CE(v[3], v[11]); SAP_CE(v[4], v[12]); SAP_CE(v[5], v[13]); \
SAP_CE(v[6], v[14]); SAP_CE(v[7], v[15]); SAP_CE(v[0], v[4]); SAP_CE(v[1], v[5]); SAP_CE(v[2], v[6]); SAP_CE(v[3], v[7]); \
SAP_CE(v[8], v[12]); SAP_CE(v[9], v[13]); SAP_CE(v[10], v[14]); SAP_CE(v[11], v[15]); SAP_CE(v[0], v[2]); SAP_CE(v[1], v[3]); \
SAP_CE(v[4], v[6]); SAP_CE(v[5], v[7]); SAP_CE(v[8], v[10]); SAP_CE(v[9], v[11]); SAP_CE(v[12], v[14]); SAP_CE(v[13], v[15]); \
SAP_CE(v[0], v[1]); SAP_CE(v[2], v[3]); SAP_CE(v[4], v[5]); SAP_CE(v[6], v[7]); SAP_CE(v[8], v[9]); SAP_CE(v[10], v[11]); \
SAP_CE(v[12], v[13]); SAP_CE(v[14], v[15]);
