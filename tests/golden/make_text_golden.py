"""Generates tests/golden/text_golden.zip: the on-disk formats either side of the hot path, written by the
UNMODIFIED reference (lib/grid.h:448-503 write, :509-674 multi_write / LAMMPS table,
:712-835 read; lib/edm_bias.cpp:586-599 HILLS lines, :166-167 + :1066-1072 restart).

    make -C oracle text            # links electronic-dance-music_b200/tests_host/text_io_driver.cpp against /root/reference/lib
    python tests/golden/make_text_golden.py

The archive holds, per case of the driver (rdf1d, coord2d, coord3d) and for its restarted second generation:
<case>_BIAS, _HIST, _HILLS_0, _MULTI, [_LMULTI]; FIX{1,2,3}.out = the reference's own fixtures read and written
back by the reference; ref_fixtures/{1,2,3}.grid = those fixtures themselves (test data of the reference's suite,
tests/1.grid ...), so that the GPU box, which has no /root/reference, can feed them to this repo's reader.
(<case>_LTAB is not stored: in the reference's serial build write_lammps_table is write_bias, lib/edm_bias.cpp:254-262;
the test compares it with <case>_BIAS.)  Runs only where /root/reference exists; the archive is committed.
"""
import os
import shutil
import subprocess
import sys
import tempfile
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_TESTS = "/root/reference/tests"
DRIVER = os.path.join(ROOT, "oracle", "_ref", "text_io_ref")


def main():
    if not os.path.isdir(REF_TESTS):
        sys.exit("the reference tree is not here: the committed archive stands")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "text"])
    tmp = tempfile.mkdtemp()
    out = subprocess.run([DRIVER, REF_TESTS], cwd=tmp, capture_output=True, text=True, check=True).stdout
    assert "TEXT_IO_DRIVER_OK" in out, out
    dst = os.path.join(ROOT, "tests", "golden", "text_golden.zip")
    with zipfile.ZipFile(dst, "w", zipfile.ZIP_DEFLATED, compresslevel=9) as z:
        for name in sorted(os.listdir(tmp)):
            if name.endswith(".edm") or name.endswith("_LTAB"):
                continue
            z.write(os.path.join(tmp, name), name)
        for d in (1, 2, 3):
            z.write(os.path.join(REF_TESTS, "%d.grid" % d), "ref_fixtures/%d.grid" % d)
        z.writestr("driver_stdout.txt", out)
    shutil.rmtree(tmp)
    print(dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
