// test helper: dumps the product's derived lookup tables (isomc_tables.h) as raw bytes on stdout
#include <cstdio>
#include "../isosurface_b200/csrc/isomc_tables.h"
int main() {
    static McTables t;
    if (isomc_build_tables(&t)) return 1;
    fwrite(&t, sizeof t, 1, stdout);
    return 0;
}
