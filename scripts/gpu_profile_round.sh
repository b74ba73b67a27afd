#!/bin/bash
# One gpurun call: GPU tests, the bench line, the ncu launch list of one bench step and `--set full` captures of the hot kernels.
# The .ncu-rep files are summarised ON THE BOX (scripts/ncu_summary.py) and deleted: gpurun only brings back 64 MiB.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_profile_round.sh r02n [notests]'
tag=${1:-r02x}
o=gpurun_out
mkdir -p $o
if [ "$2" != "notests" ]; then
  python -m pytest tests -m gpu -x -q > $o/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $o/${tag}_tests.log
fi
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; echo "bench rc=$?"
python scripts/profile_step.py > $o/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file $o/${tag}_launches.csv python scripts/profile_step.py > $o/${tag}_ncu_list.log 2>&1; echo "list rc=$?"
python scripts/profile_step.py --part matcha --steps 1 > $o/${tag}_plain1.log 2>&1 &&
ncu --set full --clock-control none --profile-from-start off -k 'regex:resnet_tc|ff_tc|attn_tc|attn_enc_tc|conv_tc' \
    -o $o/${tag}_full_matcha -f python scripts/profile_step.py --part matcha --steps 1 > $o/${tag}_ncu_full1.log 2>&1; echo "full matcha rc=$?"
python scripts/ncu_summary.py $o/${tag}_full_matcha.ncu-rep > $o/${tag}_ncu_full_matcha.txt 2>&1; rm -f $o/${tag}_full_matcha.ncu-rep
python scripts/profile_step.py --part vocoder > $o/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --profile-from-start off -k 'regex:resblock_tc|conv_tc|conv_post' \
    -o $o/${tag}_full_vocoder -f python scripts/profile_step.py --part vocoder > $o/${tag}_ncu_full2.log 2>&1; echo "full vocoder rc=$?"
python scripts/ncu_summary.py $o/${tag}_full_vocoder.ncu-rep > $o/${tag}_ncu_full_vocoder.txt 2>&1; rm -f $o/${tag}_full_vocoder.ncu-rep
python scripts/ncu_table.py matcha=$o/${tag}_ncu_full_matcha.txt vocoder=$o/${tag}_ncu_full_vocoder.txt > $o/${tag}_ncu_full_table.txt 2>&1
python scripts/ncu_launch_summary.py $o/${tag}_launches.csv $o/${tag}_launch_summary.json > $o/${tag}_launch_summary.txt 2>&1
du -sh $o; ls -la $o | tail -14
