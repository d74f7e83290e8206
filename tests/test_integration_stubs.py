"""CPU: the C stubs INTEGRATION.md shows a reference maintainer (section 2) are real code: they are cut out of the document
and compiled against the reference's own headers and include/necat_b200.h.  Needs /root/reference (this container only)."""
import os
import re
import subprocess

import pytest

import util

REF_SRC = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference sources are not on this machine")
def test_integration_stubs_compile(tmp_path):
    md = open(os.path.join(util.ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```c\n(.*?)```", md, flags=re.S)
    assert len(blocks) >= 2
    for i, code in enumerate(blocks):
        src = tmp_path / f"stub{i}.c"
        src.write_text("#include <stdio.h>\n#include <stdlib.h>\n#include <string.h>\n" + code)
        p = subprocess.run(["gcc", "-std=gnu99", "-fsyntax-only", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                            "-Werror=int-conversion", "-I", REF_SRC, "-I", os.path.join(util.ROOT, "include"), str(src)],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-2000:]


def test_option_structs_mirror_the_reference_field_for_field():
    """The stubs memcpy MapOptions / CnsOptions into nb_map_options / nb_cns_options: same size and layout."""
    if not os.path.isdir(REF_SRC):
        pytest.skip("the reference sources are not on this machine")
    code = r'''
#include <stdio.h>
#include <stddef.h>
#include "necat_b200.h"
#include "common/map_options.h"
#include "consensus/cns_options.h"
#include "common/gapped_candidate.h"
#define SAME(a, b, f) if (offsetof(a, f) != offsetof(b, f)) { printf("offset of %s differs\n", #f); return 1; }
int main(void) {
    if (sizeof(MapOptions) != sizeof(nb_map_options)) { printf("MapOptions %zu vs %zu\n", sizeof(MapOptions), sizeof(nb_map_options)); return 1; }
    if (sizeof(CnsOptions) != sizeof(nb_cns_options)) { printf("CnsOptions %zu vs %zu\n", sizeof(CnsOptions), sizeof(nb_cns_options)); return 1; }
    if (sizeof(PackedGappedCandidate) != sizeof(nb_packed_candidate)) { printf("PackedGappedCandidate\n"); return 1; }
    SAME(MapOptions, nb_map_options, kmer_size) SAME(MapOptions, nb_map_options, kmer_cnt_cutoff) SAME(MapOptions, nb_map_options, block_score_cutoff)
    SAME(MapOptions, nb_map_options, ddfs_cutoff) SAME(MapOptions, nb_map_options, error) SAME(MapOptions, nb_map_options, num_output)
    SAME(CnsOptions, nb_cns_options, min_align_size) SAME(CnsOptions, nb_cns_options, min_cov) SAME(CnsOptions, nb_cns_options, error)
    SAME(CnsOptions, nb_cns_options, mapping_ratio) SAME(CnsOptions, nb_cns_options, rescue_long_indels)
    printf("OK\n");
    return 0;
}
'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "layout.c")
        open(src, "w").write(code)
        exe = os.path.join(td, "layout")
        p = subprocess.run(["gcc", "-std=gnu99", "-I", REF_SRC, "-I", os.path.join(util.ROOT, "include"), src, "-o", exe], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-2000:]
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.stdout.strip() == "OK", out.stdout
