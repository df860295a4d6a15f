#!/usr/bin/env bash
# Entry point kept from the reference (src/train_vilt.sh:1-19): same flags, same main module role.
# `accelerate launch --config_file accelerate_config.yaml` becomes torchrun (accelerate is not in
# this image; one process per GPU, one federated client per GPU).  NGPU defaults to 1.
NGPU=${NGPU:-1}
TOKENIZERS_PARALLELISM=false python -m torch.distributed.run --nnodes=1 --nproc-per-node "${NGPU}" \
--master-addr 127.0.0.1 --master-port "${MASTER_PORT:-6012}" \
-m feddat_b200.train.main \
--encoder_name vilt \
--pretrained_model_name ./models/vilt-b32-mlm \
--climb_data_dir ''  \
--do_train  \
--model_path ./models \
--output_dir ./logs  \
--batch_size 2 \
--val_batch_size 2 \
--comm_round 30 \
--local_epochs 1 \
--lr 1e-4  \
--optimizer_mode dat \
--seed 1 \
--adapter_reduction_factor 16 \
--adapter_config pfeiffer \
--splits train_small val test_small \
--ordered_cl_tasks domain "$@"
