"""Trainer — drop-in for the reference `trainer.Trainer` (trainer.py:38-284) on the B200 backend.

Same constructor `Trainer(train_data_loader, val_data_loader, config)`, same methods (`build_model`, `train_val`,
`test`, `save_network`, `load_network`, `reset_grad`, `denorm`) and the same hot loop
(trainer.py:165-176): forward -> zero_grad -> loss -> backward -> optimiser step, with the every-10th-iteration
statistics block (trainer.py:177-218).  Differences, all additive:
  * model / loss / optimiser / metrics are the libclk-backed drop-ins of this package;
  * `config.old_model_path` (+ `distill_T`, `distill_lambda`, `num_old_classes`) turns on the continual-learning
    distillation term against a frozen previous-task network;
  * under torchrun (WORLD_SIZE > 1) every rank trains on its shard and gradients are all-reduced over NCCL
    (the reference used single-process nn.DataParallel, trainer.py:120-122); rank 0 prints and checkpoints;
  * host-side breakages of the reference on current library versions are not reproduced
    (`save_image(range=...)`, trainer.py:195; `raise ('...')`, trainer.py:92).
"""
import os
import time
from datetime import timedelta

import torch
from torch.optim.lr_scheduler import LambdaLR

from . import metrics as mt
from . import parallel
from . import voc
from .loss import CrossEntropyDistillLoss
from .optim import FusedAdam
from .unet import UNet

def to_rgb(labels):
    """label map [B,H,W] -> colour image [B,3,H,W] in [0,1]: the device form of datasets/voc.py:74-89
    (`voc.to_rgb`, float64 0..224 like the reference's) scaled for `save_image`."""
    return (voc.to_rgb(labels.contiguous()) / 255.0).float()


class Trainer:
    def __init__(self, train_data_loader, val_data_loader, config):
        self.cfg = config
        self.train_data_loader = train_data_loader
        self.val_data_loader = val_data_loader
        self.rank, self.local_rank, self.world = parallel.init_from_env()
        if not torch.cuda.is_available():
            raise RuntimeError("the B200 backend needs a CUDA device (sm_100a); there is no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.start_epoch = 0
        self.build_model()

    # ------------------------------------------------------------------ helpers (trainer.py:54-64)
    def denorm(self, x):
        return ((x + 1) / 2).clamp_(0, 1)

    def reset_grad(self):
        self.optim.zero_grad()

    # ------------------------------------------------------------------ checkpoints (trainer.py:68-102)
    def save_network(self, network, network_label, epoch_label, gpu_ids, epoch, optimizer, scheduler):
        if self.rank != 0:
            return
        path = os.path.join(self.cfg.model_save_path, "%s_net_%s.pth" % (epoch_label, network_label))
        print(path)
        state = {"epoch": epoch + 1,
                 "model_state": {k: v.detach().cpu() for k, v in network.state_dict().items()},
                 "optimizer_state": optimizer.state_dict(),
                 "scheduler_state": scheduler.state_dict()}
        torch.save(state, path)

    def load_network(self, network, network_label, epoch_label, epoch, optimizer, scheduler, save_dir=""):
        path = os.path.join(self.cfg.model_save_path, "%s_net_%s.pth" % (epoch_label, network_label))
        if not os.path.isfile(path):
            print("%s not exists yet!" % path)
            return False
        ck = torch.load(path, map_location="cpu")
        state = {k[len("module."):] if k.startswith("module.") else k: v for k, v in ck["model_state"].items()}
        network.load_state_dict(state)
        self.start_epoch = ck["epoch"]
        optimizer.load_state_dict(ck["optimizer_state"])
        scheduler.load_state_dict(ck["scheduler_state"])
        print("Load model Done!")
        return True

    # ------------------------------------------------------------------ model builder (trainer.py:105-129)
    def build_model(self):
        cfg = self.cfg
        torch.manual_seed(getattr(cfg, "seed", 0))  # identical initial weights on every rank
        self.model = UNet(num_classes=21, in_dim=3, conv_dim=64)
        self.optim = FusedAdam(self.model.parameters(), lr=cfg.lr, betas=[cfg.beta1, cfg.beta2])
        self.scheduler = LambdaLR(self.optim, lr_lambda=lambda n_iter: (1 - n_iter / cfg.n_iters) ** cfg.lr_exp)
        old = None
        if getattr(cfg, "old_model_path", None):
            old = UNet(num_classes=cfg.num_old_classes, in_dim=3, conv_dim=64)
            ck = torch.load(cfg.old_model_path, map_location="cpu")
            old.load_state_dict(ck.get("model_state", ck))
            old = old.to(self.device).eval()
            for p in old.parameters():
                p.requires_grad_(False)
        self.c_loss = CrossEntropyDistillLoss(old, T=getattr(cfg, "distill_T", 2.0), lam=getattr(cfg, "distill_lambda", 1.0))
        if cfg.continue_train:
            self.load_network(self.model, "UNET_VOC", cfg.which_epoch, self.start_epoch, self.optim, self.scheduler)
        self.model = self.model.to(self.device)
        for state in self.optim.state.values():
            for k, v in state.items():
                if torch.is_tensor(v) and v.dim() > 0:
                    state[k] = v.to(self.device)
        self.n_gpu = self.world
        if self.world > 1:
            print("Use data parallel model(# gpu: {})".format(self.world))
            parallel.attach(self.model, self.optim)

    # ------------------------------------------------------------------ training (trainer.py:132-265)
    def train_val(self):
        cfg = self.cfg
        since = time.time()
        iters_per_epoch = len(self.train_data_loader.dataset) // cfg.train_batch_size
        epoch = self.start_epoch
        if self.rank == 0:
            print(f"batch size {cfg.train_batch_size} dataset size : [{len(self.train_data_loader.dataset)}]"
                  f" epoch : [{cfg.n_iters}] iterations per epoch: {iters_per_epoch}")
        while epoch < cfg.n_iters:
            if self.rank == 0:
                print("Epoch {}/{}".format(epoch, cfg.n_iters))
                print("-" * 10)
            self.scheduler.step()  # stepped at epoch start, before any optimiser step (trainer.py:147)
            running_loss, running_corrects, total_train = 0.0, 0, 0.0
            pixel_accuracy_epoch, print_number = 0.0, 0
            start_epoch = time.time()
            for I, (input_images, target_masks) in enumerate(self.train_data_loader):
                start_mini_batch = time.time()
                inputs = input_images.to(self.device, non_blocking=True)
                labels = target_masks.to(self.device, non_blocking=True)
                if self.world > 1:
                    inputs = parallel.shard_batch(inputs, self.rank, self.world)
                    labels = parallel.shard_batch(labels, self.rank, self.world)
                self.c_loss.observe(inputs)          # frozen old-model forward (no-op without distillation)
                outputs = self.model(inputs)         # trainer.py:172
                self.reset_grad()                    # trainer.py:173
                loss = self.c_loss(outputs, labels)  # trainer.py:174
                loss.backward()                      # trainer.py:175
                self.optim.step()                    # trainer.py:176
                if I % 10 == 0:
                    print_number += 1
                    curr_loss = loss.item()
                    self.c_loss.check_labels()  # a label outside [0, 21) raises like nn.CrossEntropyLoss
                    running_loss += curr_loss
                    # argmax(softmax(x)) == argmax(x): one fused pass gives predictions, correct count and the
                    # 22x22 confusion matrix of trainer.py:183-188
                    output_label, correct, conf = mt.predict_and_count(outputs.detach(), labels, 22)
                    running_corrects += int(correct.item())
                    pixel_accuracy, total_train, _ = mt.pixel_acc(labels, output_label, total_train, running_corrects)
                    pixel_accuracy_epoch += pixel_accuracy
                    pixel_acc, pixel_acc_class, mean_IU_2, max_per_class_acc = mt.metrics_from_matrix(conf)
                    mean = mt.mean_IU_(labels, output_label)
                    self._dump_samples(inputs, labels, output_label, epoch, I)
                    if self.rank == 0:
                        elapsed = str(timedelta(seconds=time.time() - start_mini_batch))
                        print(f"Iteration : [{epoch}/{cfg.n_iters}]\tminibatch: [{I}/{iters_per_epoch}]\t"
                              f"Mini Batch Time : {elapsed}\tPixel Accuracy : {pixel_accuracy:.4f}\t"
                              f"Pixel ACC2 : {float(pixel_acc):.4f}\tPixel MAX CLASS : {float(max_per_class_acc):.4f}\t"
                              f"Class Accuracy : {float(pixel_acc_class):.4f}\tMean  : {float(mean):.4f}\t"
                              f"Mean  : {float(mean_IU_2):.4f}\tMini Batch Loss : {curr_loss:.4f}\t")
            if (epoch + 1) % cfg.log_step == 0 and self.rank == 0 and print_number:
                print(f"Iteration : [{epoch}/{cfg.n_iters}]\tEpoch Time : {timedelta(seconds=time.time() - start_epoch)}\t"
                      f"Total Time : {timedelta(seconds=time.time() - since)}\t"
                      f"Accuracy Epoch : {pixel_accuracy_epoch / print_number}\t"
                      f"Loss Epoch: {running_loss / print_number:.4f}\t")
            if (epoch + 1) % 150 == 0:
                test_acc = self.test()
                if self.rank == 0:
                    print(f"Iteration : [{epoch}/{cfg.n_iters}]\tTest Accuracy  : {test_acc}\t")
            epoch += 1
            self.save_network(self.model, "UNET_VOC", "latest", [0], epoch, self.optim, self.scheduler)
            if epoch % 10 == 0:
                self.save_network(self.model, "UNET_VOC", f"{epoch}", [0], epoch, self.optim, self.scheduler)
        if self.rank == 0:
            t = time.time() - since
            print("Training complete in {:.0f}m {:.0f}s".format(t // 60, t % 60))

    def _dump_samples(self, inputs, labels, output_label, epoch, I):
        """the three JPEG strips of trainer.py:193-195 (rank 0 only; skipped when torchvision is unavailable)."""
        if self.rank != 0 or not getattr(self.cfg, "sample_save_path", None):
            return
        try:
            import torchvision as tv
        except ImportError:
            return
        base = self.cfg.sample_save_path
        for sub in ("generated", "ground_truth", "inputs"):
            os.makedirs(os.path.join(base, sub), exist_ok=True)
        tv.utils.save_image(to_rgb(output_label).cpu(), os.path.join(base, "generated", f"predicted_{epoch}_{I}.jpg"))
        tv.utils.save_image(to_rgb(labels).cpu(), os.path.join(base, "ground_truth", f"ground_truth_{epoch}_{I}.jpg"))
        tv.utils.save_image(inputs.cpu(), os.path.join(base, "inputs", f"input_{epoch}_{I}.jpg"), normalize=True,
                            value_range=(-1, 1))

    # ------------------------------------------------------------------ eval (trainer.py:270-284)
    def test(self):
        """pixel accuracy over the validation loader; like the reference, the model is left in eval mode."""
        self.model.eval()
        parallel.broadcast_buffers(self.model)  # every rank evaluates with rank 0's running statistics
        correct = torch.zeros(1, device=self.device, dtype=torch.int64)
        total = 0
        with torch.no_grad():
            for images, labels in self.val_data_loader:
                images, labels = images.to(self.device), labels.to(self.device)
                if self.world > 1:  # no drop_last on the val loader (main.py:39-41): the last batch may be ragged
                    images = parallel.shard_batch(images, self.rank, self.world, ragged=True)
                    labels = parallel.shard_batch(labels, self.rank, self.world, ragged=True)
                if images.shape[0] == 0:
                    continue
                # forward + head + argmax + correct count in inference mode, logits never materialised
                self.model.evaluate_batch(images, labels, correct=correct)
                total += labels.nelement()
        tot = torch.tensor([total], device=self.device, dtype=torch.int64)
        _, both = parallel.all_reduce_confusion(correct.new_zeros(0), torch.cat([correct, tot]))
        return 100 * float(both[0]) / float(both[1])
