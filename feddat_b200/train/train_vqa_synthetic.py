"""Per-client VQA trainer (mirror of reference src/train/visionlanguage_tasks/train_vqa_crossvqa.py,
``VQATrainerCross``): loaders + hyper-parameters around ``TaskTrainer``.  The reference's datasets
live on the authors' NFS and are not on this box, so the loaders stream seeded synthetic batches of
the real shapes -- the batch CONTRACT (pre-encoded ``encodings`` +
``target_scores``) is what the hot path consumes."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..modeling.albef import convert_batch_to_albef_input_dict
from ..modeling.vilt import convert_batch_to_vilt_input_dict
from ..synthetic import SyntheticAlbefLoader, SyntheticVQALoader, make_albef_eval_set
from .task_trainer import TaskTrainer


class VQATrainerSynthetic(TaskTrainer):
    def __init__(self, logger, args, task_configs, model_config, device, task_key, task_output_dir=None,
                 client_id=-1, accelerator=None):
        super().__init__()
        self.accelerator = accelerator
        self.device = self.accelerator.device
        self.logger, self.args = logger, args
        self.task_key, self.task_output_dir = task_key, task_output_dir
        self.vqa_config = task_configs[task_key]
        self.albef = "albef" in args.encoder_name
        self.local_epochs = args.local_epochs
        cid = client_id if client_id >= 0 else args.ordered_cl_tasks.index(task_key)
        if self.albef:
            self.batch2inputs_converter = convert_batch_to_albef_input_dict
            self.vqa_train_dataloader = SyntheticAlbefLoader(args.synthetic_batches, args.batch_size, self.device,
                                                             image_size=args.image_size, seed=args.seed * 131 + cid,
                                                             client=cid % 8)
            self._eval_set = make_albef_eval_set(max(1, args.synthetic_batches // 4), args.val_batch_size,
                                                 args.image_size, seed=args.seed * 131 + cid + 5000)
            self.vqa_val_dataloader = self.vqa_test_dataloader = self._eval_set[2]
        else:
            self.batch2inputs_converter = convert_batch_to_vilt_input_dict
            common = dict(batch_size=args.batch_size, device=self.device, text_len=args.text_len,
                          image_size=args.image_size, num_labels=self.vqa_config["num_labels"], client=cid % 8)
            self.vqa_train_dataloader = SyntheticVQALoader(args.synthetic_batches, seed=args.seed * 131 + cid, **common)
            common["batch_size"] = args.val_batch_size
            self.vqa_val_dataloader = SyntheticVQALoader(max(1, args.synthetic_batches // 4),
                                                         seed=args.seed * 131 + cid + 5000, **common)
            self.vqa_test_dataloader = self.vqa_val_dataloader
        # train_vqa_crossvqa.py:233-239
        self.num_epochs = args.num_epochs
        self.lr = args.lr
        self.adam_epsilon = self.vqa_config["adam_epsilon"]
        self.weight_decay = self.vqa_config["weight_decay"]
        self.loss_criterion = nn.BCEWithLogitsLoss(reduction="mean")
        self.max_steps = len(self.vqa_train_dataloader) * self.num_epochs
        self.warmup_ratio = 0.1
        self.kl_temp = getattr(args, "kl_temp", 3.0)

    def compute_score_with_logits(self, logits, labels):
        """train_vqa_crossvqa.py:241-257."""
        idx = torch.max(logits, 1)[1].data
        one_hots = torch.zeros(*labels.size(), device=labels.device)
        one_hots.scatter_(1, idx.view(-1, 1), 1)
        return one_hots * labels

    def add_alpha(self, epoch, batch, step):
        """train_vqa_crossvqa.py:265-271 (alpha only feeds the momentum-distilled variant)."""
        alpha = 0.4 if epoch > 0 else 0.4 * min(1, step / len(self.vqa_train_dataloader))
        if isinstance(batch, dict):
            batch["alpha"] = alpha
        else:
            batch.append(alpha)
        return batch

    def _eval_albef(self, model, loader):
        """task_trainer.py:157-205: rank_answer over the client's answer list (top-k = 64), exact match."""
        ans_ids, ans_mask, _ = self._eval_set
        score, seen = 0, 0
        for images, q_ids, q_mask, gts in loader:
            batch = {"images": images, "question_ids": q_ids, "question_mask": q_mask, "answer_list_ids": ans_ids,
                     "answer_list_mask": ans_mask, "train": False, "k": min(64, ans_ids.shape[0])}
            with torch.no_grad():
                topk_ids, topk_probs = model(self.task_key, batch)
            pred = topk_ids.gather(1, topk_probs.argmax(dim=1, keepdim=True))           # [B, 1]
            score += int((pred.cpu() == gts).any(dim=1).sum())
            seen += gts.shape[0]
        return score / max(seen, 1) * 100.0

    def eval_one_loader(self, model, loader):
        """task_trainer.py:113-209; score sum / count are reduced across ranks by the caller instead of
        gathering logits."""
        model.eval()
        if self.albef:
            s = self._eval_albef(model, loader)
            model.train()
            return s
        score, seen = 0.0, 0
        for batch in loader:
            if hasattr(model, "new_step"):
                model.new_step()
            logits = self.forward_pass(model, batch, do_eval=True)[1]
            target = batch["target_scores"].to(self.device)
            score += torch.sum(self.compute_score_with_logits(logits.float(), target)).item()
            seen += target.shape[0]
        model.train()
        return score / max(seen, 1) * 100.0

    def eval(self, model):
        """task_trainer.py:211-246: [gating, adapter_0 alone, adapter_1 alone].  Leaves
        ``adapter_0.requires_grad = False`` behind exactly like the reference."""
        loader = self.vqa_val_dataloader if "gqa" in self.task_key else self.vqa_test_dataloader
        model.activate_gating()
        s = self.eval_one_loader(model, loader)
        model.deactivate_gating()
        model.set_active_adapter("adapter_0")
        s0 = self.eval_one_loader(model, loader)
        model.deactivate_gating()
        model.set_active_adapter("adapter_1")
        s1 = self.eval_one_loader(model, loader)
        return [s, s0, s1]
