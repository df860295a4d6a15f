#!/usr/bin/env bash
# Entry point kept from the reference (src/train_albef.sh:1-18): same flags, same main module role.
# `accelerate launch --config_file accelerate_config.yaml` becomes torchrun (accelerate is not in this
# image; one process per GPU, one federated client per GPU).  NGPU defaults to 1.  Without ./models/ALBEF.pth
# on disk (there is no network on the box) the architecture is built with seeded random weights.
NGPU=${NGPU:-1}
TOKENIZERS_PARALLELISM=false python -m torch.distributed.run --nnodes=1 --nproc-per-node "${NGPU}" \
--master-addr 127.0.0.1 --master-port "${MASTER_PORT:-6013}" \
-m feddat_b200.train.main \
--encoder_name albef_no_distill \
--pretrained_model_name ./models/ALBEF.pth \
--climb_data_dir ''  \
--do_train  \
--model_path ./models/ \
--output_dir ./logs/  \
--batch_size 2 \
--val_batch_size 2 \
--lr 1e-4  \
--optimizer_mode dat \
--seed 2 \
--adapter_reduction_factor 16 \
--adapter_config pfeiffer \
--splits train_small val test \
--ordered_cl_tasks domain "$@"
