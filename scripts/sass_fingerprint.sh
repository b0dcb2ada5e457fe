# Fingerprint of the product's render kernels (pbrtgpu.cu's cubin inside libpbrtgpu.so).  The kernels that ran the round-2
# final GPU visit (profiles/r02g_*: 94 GPU tests, 350.9 Msamples/s) have the fingerprint below -- checked at the end of the
# round by rebuilding that visit's commit (87e4ae1) next to HEAD: same SASS -- and a change to an experiment that is supposed
# to be compiled out (PBRT_METAL, PBRT_PAIR_NODES, PBRT_TRACE_LOOP ...) must leave it untouched.   bash scripts/sass_fingerprint.sh
set -e
want=95c852e91f9dcf274cde533f38560f3b
tmp=$(mktemp -d)
( cd $tmp && cuobjdump -xelf all "$OLDPWD/pbrt_v2_b200/libpbrtgpu.so" > /dev/null )
got=$(cuobjdump -sass $tmp/pbrtgpu.sm_100a.cubin | md5sum | cut -d' ' -f1)
rm -rf $tmp
echo "pbrtgpu.sm_100a.cubin sass md5: $got (round-2 GPU-verified kernels: $want)"
[ "$got" = "$want" ]
