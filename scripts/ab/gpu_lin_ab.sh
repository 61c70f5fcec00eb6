for so in "" $PWD/crog_b200/lib/libcrog_b200.epimin.so; do
echo "lib [$so]"
for shape in "43264 512 512 0" "43264 512 512 5" "43264 1536 512 5" "43264 2048 512 5" "43264 512 2048 0" "1088 1536 512 0" "1088 512 2048 0"; do CROG_B200_SO=$so python tests/prof_linear.py $shape 2>&1 | tail -1; done
done
