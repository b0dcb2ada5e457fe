# Fingerprint of the product's render kernels (pbrtgpu.cu's cubin inside libpbrtgpu.so).  The kernels that ran the
# round-1 GPU parity suite and produced profiles/r01_* have the fingerprint below; a change to an experiment that is
# supposed to be compiled out (PBRT_SPECULAR, PBRT_PAIR_NODES ...) must leave it untouched.   bash scripts/sass_fingerprint.sh
set -e
want=8e08f91a4c85b9de6b8ae35b1177288e
tmp=$(mktemp -d)
( cd $tmp && cuobjdump -xelf all "$OLDPWD/pbrt_v2_b200/libpbrtgpu.so" > /dev/null )
got=$(cuobjdump -sass $tmp/pbrtgpu.sm_100a.cubin | md5sum | cut -d' ' -f1)
rm -rf $tmp
echo "pbrtgpu.sm_100a.cubin sass md5: $got (round-1 GPU-verified kernels: $want)"
[ "$got" = "$want" ]
