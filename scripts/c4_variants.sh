for lib in build/lib_*.so; do echo "== $lib"; PYPORE_B200_LIB=$PWD/$lib timeout 200 python scripts/config_times.py 2>&1 | grep -E "C4" | cut -c1-140; done
