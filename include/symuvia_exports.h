/* symuvia_exports.h — the SymuVia C exports re-implemented by libSymuViaB200.so.
 *
 * Same names, signatures, ownership and error conventions as the reference's SymExpC.h:15-53 and their
 * definitions in symuvia/src/Symubruit.cpp:4239-4598, so that a caller binding libSymuVia.so (Python/ctypes,
 * SymuMaster, EVE) can bind this library instead.  Each entry cites the reference definition it replaces.
 */
#ifndef SYMUVIA_EXPORTS_H
#define SYMUVIA_EXPORTS_H
#ifdef __cplusplus
extern "C" {
#endif

/* Symubruit.cpp:4241-4264 — load, run every step, unload. 0 = ok, 1 = load failure. */
int SymRunEx(char* sFile);
/* Symubruit.cpp:4266-4269 — returns the bool load result as int (1 ok, 0 failure). Accepts the reference's
 * ROOT_SYMUBRUIT XML (the subset read by csrc/xmlnet.h, listed in INTEGRATION.md) or an already flattened *.symflat file. */
int SymLoadNetworkEx(char* sFile);
/* Symubruit.cpp:4271-4274 */
void SymUnloadCurrentNetworkEx(void);
/* Symubruit.cpp:4276-4279 -> :992-1016: one time step; the <INST> fragment is strcpy'd into the caller-allocated
 * buffer (no length check, as in the reference); *bNEnd = true when start + t >= end (reseau.cpp:2427), and the
 * networks are then released (Symubruit.cpp:1010-1011). false when no network is loaded. */
bool SymRunNextStepEx(char* sOutputXmlFlow, bool bTrace, bool* bNEnd);
/* Symubruit.cpp:4281-4284 -> :1051 — same without building the XML text. */
bool SymRunNextStepLiteEx(bool bTrace, bool* bNEnd);
/* Symubruit.cpp:4345-4348 -> :1714-1800 — queues a vehicle for the next step; returns its id, or a negative
 * code: -1 no network, -2 unknown type, -3 unknown origin, -4 unknown destination, -5 bad lane. */
int SymCreateVehicleEx(char* sType, char* sOrigin, char* sDestination, int nLane, double dbt);
/* Symubruit.cpp:2303 — releases strings returned by the library. */
void SymRunDeleteStr(char* str);
/* Not a SymuVia export: flattens a ROOT_SYMUBRUIT document into a SYMFLAT1 file (host only, no CUDA). 0 = ok. */
int symflat_from_xml(const char* xml_path, const char* out_symflat_path);

#ifdef __cplusplus
}
#endif
#endif
