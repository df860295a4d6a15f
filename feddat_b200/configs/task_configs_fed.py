"""Per-client task configs (reference src/configs/task_configs_fed.py:162-178,256-282): the five
"domain" VQA clients of train_vilt.sh plus synthetic heterogeneous clients for the no-dataset box."""


def _vqa_client(name):
    return {
        "task_name": name,
        "splits": ["train", "val"],
        "num_labels": 100,
        "num_images": 1,
        "model_type": "classification",
        "num_epochs": 20,
        "lr": 1e-4,
        "weight_decay": 1e-2,
        "adam_epsilon": 1e-8,
        "warmup_ratio": 0.1,
        "random_baseline_score": 0.0,
    }


DOMAIN_TASKS = ["art", "abstract", "vizwiz", "toronto", "gqa"]          # main.py:358-359 ("domain")
task_configs = {n: _vqa_client(n) for n in DOMAIN_TASKS}
for _i in range(8):
    task_configs[f"synth{_i}"] = _vqa_client(f"synth{_i}")
