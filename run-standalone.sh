#!/bin/bash
# Same interface as the reference's run-standalone.sh.
. ./shared-args.sh
cd distributed-gan_b200

seed=1

python standalone_gan.py --local_epochs $local_epochs \
    --epochs $epochs \
    --model $model \
    --dataset $dataset \
    --generator_lr $generator_lr \
    --discriminator_lr $discriminator_lr \
    --device $device \
    --batch_size $batch_size \
    --seed $seed \
    --beta_1 $beta_1 \
    --beta_2 $beta_2 \
    --log_interval $log_interval
