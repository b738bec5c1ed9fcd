echo "128->128 @256 default (sliced fold)"; python tools/prof_conv_time.py 32 128 128 256 3 1 1
echo "no slices (halo<128>)"; VSP_NO_FOLD_SLICES=1 python tools/prof_conv_time.py 32 128 128 256 3 1 1
echo "no slices, no halo (pair<128>)"; VSP_NO_FOLD_SLICES=1 VSP_NO_HALO=1 python tools/prof_conv_time.py 32 128 128 256 3 1 1
echo "64->64 @512 default (ring<64,9>)"; python tools/prof_conv_time.py 32 64 64 512 3 1 1
echo "no ring (halo<64>)"; VSP_NO_RING=1 python tools/prof_conv_time.py 32 64 64 512 3 1 1
echo "no ring no halo (generic<64>)"; VSP_NO_RING=1 VSP_NO_HALO=1 python tools/prof_conv_time.py 32 64 64 512 3 1 1
echo "256->64 d2 @128 default (halo<64>)"; python tools/prof_conv_time.py 32 256 64 128 3 2 1
echo "no halo (generic<64>)"; VSP_NO_HALO=1 python tools/prof_conv_time.py 32 256 64 128 3 2 1
