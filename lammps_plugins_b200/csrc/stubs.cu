// temporary: entry points not implemented yet fail loudly
#include "common.cuh"
#define NOTIMPL(c, name) do { if (c) (c)->fail(name ": not implemented yet"); return B200MD_ERR_ARG; } while (0)
void b200md_system_free(b200md_ctx *) {}
extern "C" {
int b200md_neigh_build(b200md_ctx *c, const b200md_box *, int, const double *, const double *, int, int, const double *, const int *, int, double) { NOTIMPL(c, "neigh_build"); }
int b200md_system_create(b200md_ctx *c, const b200md_system_desc *, int, const double *, const double *, const int *, const int *) { NOTIMPL(c, "system_create"); }
int b200md_nccl_unique_id(void *) { return B200MD_ERR_ARG; }
int b200md_system_comm_init(b200md_ctx *c, const void *, int, int) { NOTIMPL(c, "system_comm_init"); }
int b200md_system_run(b200md_ctx *c, int, int) { NOTIMPL(c, "system_run"); }
int b200md_system_thermo(b200md_ctx *c, double *) { NOTIMPL(c, "system_thermo"); }
int b200md_system_thermo_count(b200md_ctx *c) { NOTIMPL(c, "system_thermo_count"); }
int b200md_system_thermo_row(b200md_ctx *c, int, double *) { NOTIMPL(c, "system_thermo_row"); }
int b200md_system_sizes(b200md_ctx *c, long long *) { NOTIMPL(c, "system_sizes"); }
int b200md_system_download(b200md_ctx *c, double *, double *, double *, int *, int *) { NOTIMPL(c, "system_download"); }
}
