# The kernels-on-the-host harnesses under AddressSanitizer + UBSan: "device" buffers are malloc'd there, so an
# out-of-bounds index in a kernel is a heap overflow the sanitizer reports.  CPU only:  bash scripts/sanitize_harnesses.sh
set -e
out=${TMPDIR:-/tmp}/pbrt_san
mkdir -p $out
FLAGS="-O1 -g -std=c++17 -ffp-contract=off -fsanitize=address,undefined -fno-sanitize-recover=undefined -I/usr/local/cuda/include"
g++ $FLAGS tests/native/bvhbuild_on_host.cpp tests/native/san/main_bvhbuild.cpp -o $out/san_bvhbuild
g++ $FLAGS tests/native/loopsubdiv_on_host.cpp tests/native/san/main_loopsubdiv.cpp -o $out/san_loopsubdiv
g++ $FLAGS -pthread tests/native/tables_on_host.cpp tests/native/san/main_tables.cpp -o $out/san_tables
$out/san_bvhbuild
$out/san_loopsubdiv
$out/san_tables
