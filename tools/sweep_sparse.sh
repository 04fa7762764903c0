# lanes vs rows on config-E-shaped sparse lists of growing size (development aid)
for cfg in ${SWEEP:-1000:1e5 2000:2e5 4000:8e5 10000:4e6}; do n=${cfg%%:*}; m=${cfg##*:}
for p in ${PATHS:-lanes rows}; do
echo "N=$n M=$m path=$p: $(E_N=$n E_M=$m E_S=${E_S:-64} E_R=${E_R:-8} EHIC_PATH=$p timeout 200 python tools/bench_e.py 2>&1 | grep 'config E')"
done; done
