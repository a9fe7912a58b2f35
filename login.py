"""Log this host into Weights & Biases (reference ``login.py:9-22`` logs every TPU host in).
``python login.py --key <API key>``; with no network the key is only stored for offline runs."""
import argparse


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--key", required=True)
    a = p.parse_args(argv)
    try:
        import wandb

        wandb.login(key=a.key, relogin=True)
        print("wandb login stored")
    except Exception as e:  # offline image: keep going, metrics still go to metrics.jsonl
        print(f"wandb login skipped: {e}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
