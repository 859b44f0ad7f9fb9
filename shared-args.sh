#!/bin/bash
# Canonical hyper-parameters of the reference (shared-args.sh:3-15); device is cuda here (no CPU path).
# Set MDGAN_SYNTH_M=<samples> to train on synthetic images when the torchvision files are not on disk.

batch_size=10
discriminator_lr=0.0002
generator_lr=0.0002
dataset=CIFAR10
model=$dataset
epochs=30000
local_epochs=1
iid=1
n_samples_fid=10
device=cuda
log_interval=300
beta_1=0.5
beta_2=0.999
