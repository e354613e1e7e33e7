"""include/hp_b200.h must be plain C (the drop-in boundary is a C ABI): compile a C99 translation unit that includes it,
references every declared function and calls the host-only ones against libhp_b200.so."""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "hp_b200.h")


def test_header_compiles_as_c99_and_links(hp, tmp_path):
    lib_dir = os.path.dirname(hp._native.LIB_PATH)
    src = open(HEADER).read()
    names = sorted(set(re.findall(r"HP_API\s+[\w\s\*]+?\b(hp_\w+)\s*\(", src)))
    assert len(names) >= 20
    refs = "\n".join(f"    p[{i}] = (void (*)(void))&{n};" for i, n in enumerate(names))
    c = tmp_path / "abi.c"
    c.write_text(f"""
#include <stdio.h>
#include <string.h>
#include "hp_b200.h"
int main(void) {{
    void (*p[{len(names)}])(void);
{refs}
    int dims[6] = {{3, 32, 64, 128, 64, 3}};
    if (hp_version() != HP_B200_VERSION) return 1;
    if (strcmp(hp_error_string(HP_ERR_WORKSPACE), "workspace too small") != 0) return 2;
    if (hp_target_network_num_weights(5, dims, 1) != 19011) return 3;
    if (hp_chamfer_workspace_bytes(32, 2048, 2048) < (size_t)8 * 32 * 4096) return 4;
    if (hp_nndistance(-1, 1, NULL, 1, NULL, NULL, NULL, NULL, NULL, NULL) != HP_ERR_INVALID_ARGUMENT) return 5;
    if (strstr(hp_last_error_message(), "negative size") == NULL) return 6;
    printf("%d symbols ok\\n", (int)(sizeof(p) / sizeof(p[0])));
    return p[0] == NULL;
}}
""")
    exe = tmp_path / "abi"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(REPO, "include"),
                           str(c), "-o", str(exe), "-L", lib_dir, "-lhp_b200", f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert f"{len(names)} symbols ok" in out.stdout
