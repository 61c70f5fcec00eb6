for t in "" n5 n6 n8; do
so=""; [ -n "$t" ] && so=$PWD/crog_b200/lib/libcrog_b200.$t.so
echo "variant [$t]"; CROG_B200_SO=$so python scripts/time_tail.py
done
