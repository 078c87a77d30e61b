// `recgraph` executable: thin wrapper over rg_cli_main (the same entry the parity tests call).
#include <cstdio>
#include <cstring>

#include "../../include/recgraph_b200.h"

int main(int argc, char** argv) {
    char *out = nullptr, *err = nullptr;
    int rc = rg_cli_main(argc, (const char**)argv, &out, &err);
    if (out) fwrite(out, 1, strlen(out), stdout);
    if (err) fwrite(err, 1, strlen(err), stderr);
    rg_free(out);
    rg_free(err);
    return rc;
}
