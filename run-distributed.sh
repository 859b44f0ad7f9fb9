#!/bin/bash
# Same interface as the reference's run-distributed.sh: ./run-distributed.sh <ranks>   e.g.  ./run-distributed.sh 0..2
. ./shared-args.sh
cd distributed-gan_b200

seed=3
world_size=3
backend=nccl
swap_interval=5000
master_addr=127.0.0.1
master_port=1234

python bootstrap.py \
    --backend $backend \
    --world_size $world_size \
    --dataset $dataset \
    --ranks $1 \
    --epochs $epochs \
    --local_epochs $local_epochs \
    --swap_interval $swap_interval \
    --discriminator_lr $discriminator_lr \
    --generator_lr $generator_lr \
    --model $model \
    --device $device \
    --batch_size $batch_size \
    --iid $iid \
    --seed $seed \
    --master_addr $master_addr \
    --master_port $master_port \
    --beta_1 $beta_1 \
    --beta_2 $beta_2 \
    --log_interval $log_interval &

trap "trap - SIGTERM && kill -- -$$" SIGINT SIGTERM
wait
