"""Problem layer — B200-native mirror of `mmdyn/pytorch/problems/problems.py`.

Same classes (`Problem` > `Reconstruction` > `SeqModeling` > `DynModeling`), constructor
(`Problem(args, log_dir=None, load_dataset=None)`), public methods (`train`, `parse_input`,
`_evaluate_model(x, targets, reduction, reduce) -> (outputs, loss)`, `_anneal_KL`, properties) and
on-disk artefacts (run dirs, tensorboard scalars, best-loss checkpoints with the reference's
state_dict keys) — but `_evaluate_model` / `_evaluate_mvae` hand the whole sub-sampled step to
`mmdyn_b200.engine.StepEngine` (one fused launch sequence: shared encoder trunks, group-batched
decoders, fused PoE/KL/losses and a hand-written backward), and `set_optimizer` installs the
single-kernel `FusedAdam` / `FusedSGD` over the flat parameter arena.

Differences that are deliberate and documented in DESIGN.md:
  * perf_measure values stay on the device (0-dim tensors) so the step has no host sync; they
    are converted with float() only where the reference logs them per epoch;
  * `--problem-type regression` and class-label (categorical) conditioning are outside the accelerated path;
    `--conditional` with the shock force of SeqModeling / DynModeling is supported;
  * dyn_modeling batches are sequence-collated (the reference crashes there).
"""
import math
import os
from collections import defaultdict
from datetime import datetime
from pathlib import Path

import numpy as np
import torch

from mmdyn_b200 import engine, losses, optim as fused_optim
from mmdyn_b200.pytorch import config
from mmdyn_b200.pytorch.models.models import setup_model
from mmdyn_b200.pytorch.utils.datasets import dataset_setup
from mmdyn_b200.pytorch.utils.training import progress_bar, save_pkl


def _f(v):
    return float(v) if not isinstance(v, (int, float)) else v


class Problem:
    def __init__(self, problem_args, log_dir=None, load_dataset=None):
        self._model = self._log_dir = self._checkpoint_dir = self._tensorboard_dir = self._plot_dir = None
        self._condition_dim = self._classes = self._criterion = self._optimizer = self._writer = None
        self.train_dataset = self.test_dataset = self.train_loader = self.test_loader = None
        self._best_acc, self._best_loss = 0, np.inf
        self._load_dataset = load_dataset
        self._logger_dict = defaultdict(list)
        self._logger_histogram = defaultdict(list)
        self._img_logger_dict = defaultdict()
        self._fig_logger_dict = defaultdict()
        self.parameters = vars(problem_args) if not isinstance(problem_args, dict) else dict(problem_args)
        self._cross_modal = self.parameters['input_type'] == 'visuotactile'
        self._kl_weight = self.parameters['kl_weight']
        self._pose_multiplier = self.parameters['pose_multiplier']
        self._conditional = self.parameters['conditional']
        self._categorical_conditions = None
        self._seq_length = None
        self._engine = None
        self._start_epoch = 0
        if self.parameters.get('no_cuda') or not torch.cuda.is_available():
            raise RuntimeError("mmdyn_b200 runs on a CUDA device (B200) only; --no-cuda / CPU execution is the "
                               "reference's own path")
        self._device = torch.device('cuda', torch.cuda.current_device())
        assert (self.parameters['input_type'] in config.INPUT_TYPES), "Input type is not implemented"
        if log_dir:
            self.load_dir(log_dir)
            self._load_problem()
        else:
            self.set_dir()
            self._set_problem()
            if self.parameters.get('resume'):
                self.load_checkpoint(self.parameters['resume'])

    # -- construction ---------------------------------------------------------------------------
    def _set_problem(self):
        self.set_dataset()
        self.set_model()
        self.set_criterion()
        self.set_optimizer()

    def _load_problem(self):
        if self._load_dataset:
            self.set_dataset()
            self.set_model()

    def set_model(self):
        raise NotImplementedError

    def _set_condition_dim(self):
        raise NotImplementedError

    def load_dir(self, log_dir):
        self._log_dir = log_dir
        self._checkpoint_dir = log_dir + '/checkpoint/'
        self._tensorboard_dir = log_dir + '/tensorboard/'
        self._plot_dir = log_dir + '/plot/'

    def set_dir(self):
        stamp = datetime.now().strftime("_%Y_%m_%d_%H_%M_%S")
        self.load_dir('./logs/' + self.parameters['save_name'] + stamp)
        for d in (self._log_dir, self._checkpoint_dir, self._tensorboard_dir, self._plot_dir):
            Path(d).mkdir(parents=True, exist_ok=True)

    def parse_input(self, data, target):
        it = self.parameters['input_type']
        if not isinstance(data, list):
            model_input = data.to(self._device)
        elif it == 'visuotactile':
            model_input = [data[0].to(self._device), data[1].to(self._device)]
        else:
            model_input = data[{'visual': 0, 'tactile': 1}[it]].to(self._device)
        return model_input, target.to(self._device)

    def set_dataset(self):
        self._input_size = (64, 64)
        self._n_channels = 3
        self.dataset_dict = dataset_setup(self.parameters['dataset_path'], self.parameters['problem_type'],
                                          input_size=self._input_size, batchsize=self.parameters['batchsize'],
                                          shuffle=True)
        d = self.dataset_dict
        self.train_dataset, self.test_dataset = d['train_dataset'], d['test_dataset']
        self.train_loader, self.test_loader = d['train_loader'], d['test_loader']
        self._seq_length = d['seq_length']
        if 'classes' in d:
            self._classes = d['classes']

    def set_criterion(self):
        raise NotImplementedError

    def set_optimizer(self):
        """problems.py:130-138 with the fused single-kernel updates."""
        name = self.parameters['optimizer']
        assert (name in config.OPTIMIZERS), "loss name not implemented in Problem"
        if name == 'SGD':
            self._optimizer = fused_optim.FusedSGD(self._model, lr=self.parameters['lr'], momentum=0.9,
                                                   weight_decay=5e-4)
        else:
            self._optimizer = fused_optim.FusedAdam(self._model, lr=self.parameters['lr'])

    def _evaluate_model(self, inputs, targets, **kwargs):
        raise NotImplementedError

    def _graphed_step(self, inputs, targets):
        """Optional fast path of the training step (see Reconstruction); None = run it eagerly."""
        return None

    # -- epochs (problems.py:143-216) -----------------------------------------------------------------
    def _train_epoch(self, epoch):
        print('Epoch: %d' % epoch)
        self._model.train()
        train_loss = torch.zeros((), device=self._device)
        inputs = outputs = targets = []
        perf_measure = {'visual': 0, 'tactile': 0, 'pose': 0}
        n = len(self.train_loader)
        for batch_idx, (data_input, data_target) in enumerate(self.train_loader):
            inputs, targets = self.parse_input(data_input, data_target)
            step = self._graphed_step(inputs, targets)
            if step is not None:
                # zero_grad + fused forward + fused backward + fused optimizer: one CUDA-graph replay
                outputs, loss = step
            else:
                self._optimizer.zero_grad()
                outputs, loss = self._evaluate_model(inputs, targets)
                loss.backward()
                self._optimizer.step()
            train_loss += loss.detach()
            for k, v in outputs.get('perf_measure', {}).items():
                perf_measure[k] = perf_measure[k] + v
            if batch_idx % 50 == 0 or batch_idx == n - 1:
                # the host reads the loss here anyway: read the optimizer's non-finite flag with it and fail
                # loudly instead of training on inf / NaN (fp16 operands)
                lv = loss.item()
                if hasattr(self._optimizer, 'check_finite'):
                    self._optimizer.check_finite(lv)
                if self._writer is not None:
                    self._writer.add_scalar('Loss/train_step', lv, epoch * n + batch_idx)
                    progress_bar(batch_idx + 1, n, 'Loss %.3f' % lv)
        self._log_train_info(inputs, outputs, targets, _f(train_loss), epoch, perf_measure=perf_measure)
        return perf_measure

    def _test_epoch(self, epoch):
        self._model.train()  # sic: the reference validates with batch statistics and dropout on (:174)
        validation_loss = torch.zeros((), device=self._device)
        inputs = outputs = targets = []
        perf_measure = {'visual': 0, 'tactile': 0, 'pose': 0}
        with torch.no_grad():
            for batch_idx, (data_input, data_target) in enumerate(self.test_loader):
                inputs, targets = self.parse_input(data_input, data_target)
                outputs, loss = self._evaluate_model(inputs, targets)
                validation_loss += loss
                for k, v in outputs.get('perf_measure', {}).items():
                    perf_measure[k] = perf_measure[k] + v
            self._log_test_info(inputs, outputs, targets, _f(validation_loss), epoch, perf_measure=perf_measure)
        return perf_measure

    # -- checkpoints ---------------------------------------------------------------------------------
    def save_checkpoint(self, loss, epoch):
        """The reference's checkpoint, byte-compatible (problems.py:751-757: keys model / loss / epoch), plus
        a side file `<ckpt>.optim` with the fused optimizer's moments so that a resumed run continues
        the same trajectory (the reference saves no optimizer state and cannot resume)."""
        path = self._checkpoint_dir + '/epoch_' + str(epoch) + '.ckpt'
        torch.save({'model': self._model.state_dict(), 'loss': float(loss), 'epoch': int(epoch)}, path)
        if hasattr(self._optimizer, 'state_dict') and isinstance(self._optimizer, fused_optim._FlatOptimizer):
            side = self._optimizer.state_dict()
            # device noise stream (dropout masks, eps) of the graphed step: seed + Philox counter, so that a
            # resumed run draws the numbers the uninterrupted run would have drawn next
            src = self._engine._noise() if self._engine is not None else None
            if src is not None and hasattr(src, 'state_dict'):
                side['noise'] = src.state_dict()
            torch.save(side, path + '.optim')
        return path

    def load_checkpoint(self, path):
        """Resume: model weights + BatchNorm buffers from `path` (written here or by the reference),
        best loss and epoch counter; optimizer moments from `<path>.optim` when present (otherwise the
        optimizer restarts from zero moments, which is all a reference checkpoint allows)."""
        # weights_only: a checkpoint is tensors + a float + an int; never unpickle arbitrary objects from
        # a user-supplied path (reference checkpoints, third-party files)
        state = torch.load(path, map_location='cpu', weights_only=True)
        missing = {'model', 'loss', 'epoch'} - set(state)
        if missing:
            raise ValueError(f"{path} is not a mmdyn checkpoint: missing {sorted(missing)}")
        self._model.load_state_dict(state['model'])
        engine.get_arena(self._model).bump()  # the packed fp16 operand copies must be rebuilt
        self._best_loss = _f(state['loss'])
        self._start_epoch = int(state['epoch']) + 1
        if os.path.exists(path + '.optim') and isinstance(self._optimizer, fused_optim._FlatOptimizer):
            side = torch.load(path + '.optim', map_location='cpu', weights_only=True)
            self._optimizer.load_state_dict(side)
            if side.get('noise') is not None and hasattr(self, '_get_engine'):
                from mmdyn_b200 import noise as _noise
                self._get_engine().noise_src = _noise.DeviceNoise.from_state_dict(side['noise'], self._device)
        return state

    def train(self, save=True):
        from torch.utils.tensorboard import SummaryWriter
        perf_measure = 0
        self._writer = SummaryWriter(self._tensorboard_dir)
        first = self._start_epoch
        for epoch in range(first, self.parameters['num_epochs']):
            self._anneal_KL(epoch)
            self._train_epoch(epoch)
            perf_measure = self._test_epoch(epoch)
            self._sample(n=50)
            for key in self._logger_dict:
                self._writer.add_scalar(key, self._logger_dict[key][epoch - first], epoch)
            for key in self._logger_histogram:
                self._writer.add_histogram(key, self._logger_histogram[key], global_step=epoch)
            self._write_images(epoch)
        hp = {k: v for k, v in self.parameters.items() if isinstance(v, (int, float, str, bool))}
        if isinstance(perf_measure, dict):  # no epoch ran (resumed at or past --num-epochs): nothing to report
            self._writer.add_hparams(hp, {k: _f(v) for k, v in perf_measure.items()})
        if save:
            save_pkl(dict(self._logger_dict), os.path.join(self._log_dir, 'results.pkl'))

    def _anneal_KL(self, epoch):
        """problems.py:212-216: linear warm-up over `annealing_epochs`, overriding --kl-weight."""
        ae = self.parameters['annealing_epochs']
        self._kl_weight = (epoch + 1) / ae if epoch < ae else 1

    def _sample(self, n=50):
        raise NotImplementedError

    def _log_train_info(self, inputs, outputs, targets, loss, epoch, perf_measure=None):
        raise NotImplementedError

    def _log_test_info(self, inputs, outputs, targets, loss, epoch, perf_measure=None):
        raise NotImplementedError

    def _write_images(self, epoch, n_images=100):
        raise NotImplementedError

    log_dir = property(lambda self: self._log_dir)
    model = property(lambda self: self._model)
    checkpoint_dir = property(lambda self: self._checkpoint_dir)
    plot_dir = property(lambda self: self._plot_dir)
    dataset = property(lambda self: self.test_dataset)
    num_epochs = property(lambda self: self.parameters['num_epochs'])
    input_type = property(lambda self: self.parameters['input_type'])
    condition_dim = property(lambda self: self._condition_dim)


class Regression(Problem):
    """Baseline regressing the resting pose from the first frame (problems.py:263-359).
    Eager step: trunk + out_net.0 on the tensor-core kernels, the rest of out_net and the MSE in fp32."""

    def set_model(self):
        self._set_condition_dim()
        # the reference passes `condition_dim`, which its Regressor does not accept (TypeError as shipped,
        # problems.py:272-277 vs models.py:30); the mirror's Regressor takes it as an alias of num_classes
        self._model = setup_model(self.parameters['model_name'], condition_dim=self._condition_dim, out_dim=7,
                                  conditional=self._conditional)
        self._model.to(self._device)

    def _set_condition_dim(self):
        """problems.py:283-289: the shock-force width; the reference's fallback (`len(data[0][-1])`) only makes
        sense for its pickled list datasets, here a dataset without a shock field gives 0."""
        self._categorical_conditions = False
        try:
            self._condition_dim = len(self.train_dataset.data[0][0][4])
        except Exception:
            self._condition_dim = 0

    def parse_input(self, data, target):
        """problems.py:291-316: the image of the first frame of every sequence -> the resting pose."""
        L, dev = self._seq_length, self._device
        model_input = target_output = None
        if not isinstance(data, list):
            model_input, target_output = data.to(dev), target.to(dev)
        elif len(data) == 1:
            model_input, target_output = data[0].to(dev), target[0].to(dev)
        elif self.parameters['input_type'] == 'visual':
            model_input, target_output = data[0][::L].to(dev), target[2][::L].to(dev)
        elif self.parameters['input_type'] == 'tactile':
            model_input, target_output = data[1][::L].to(dev), target[2][::L].to(dev)
        shock = data[4][::L].to(dev) if isinstance(data, list) and len(data) > 4 else None
        return {'model_input': model_input, 'shock': shock}, target_output

    def set_criterion(self):
        self._criterion = losses.mse_sum  # nn.MSELoss(reduction='sum'), one kernel (sum + gradient)

    def _evaluate_model(self, inputs, targets, **kwargs):
        """problems.py:321-332."""
        if self._conditional:
            if inputs['shock'] is None:
                raise ValueError("--conditional needs the shock force as the 5th data field (datasets.py: data[4])")
            out = self._model(inputs['model_input'], inputs['shock'])
        else:
            out = self._model(inputs['model_input'])
        loss = self._criterion(out.view(targets.size()), targets)
        with torch.no_grad():
            pose_measure = loss.detach() / targets.numel()  # F.mse_loss(reduction='mean') of the same tensors
        return {'outputs': out, 'perf_measure': {'pose': pose_measure}}, loss

    def _sample(self, n=50):
        pass

    def _log_train_info(self, inputs, outputs, targets, loss, epoch, perf_measure=None):
        self._logger_dict['Loss/train_epoch'].append(loss / len(self.train_loader))
        if perf_measure:
            for k, v in perf_measure.items():
                self._logger_dict['Perf_measure_train/' + k].append(v / len(self.train_loader))

    def _log_test_info(self, inputs, outputs, targets, loss, epoch, perf_measure=None):
        self._logger_dict['Loss/validation_epoch'].append(loss / len(self.test_loader))
        if perf_measure:
            for k, v in perf_measure.items():
                self._logger_dict['Perf_measure_validation/' + k].append(v / len(self.test_loader))
        if loss < self._best_loss:
            self.save_checkpoint(loss, epoch)
            self._best_loss = loss

    def _write_images(self, epoch, n_images=100):
        pass


class Reconstruction(Problem):
    def set_model(self):
        self._set_condition_dim()
        name = self.parameters['model_name']
        kw = dict(condition_dim=self._condition_dim, input_dim=int(np.prod(np.array(self._input_size))),
                  architecture=name.split('-')[0], conditional=self._conditional,
                  categorical_conditions=self._categorical_conditions,
                  latent_size=self.parameters.get('latent_size', 256))
        if 'mvae' in name:
            kw['use_pose'] = self.parameters['use_pose']
        self._model = setup_model(name, cross_modal=self._cross_modal, **kw)
        self._model.to(self._device)
        self._engine = None

    def _set_condition_dim(self):
        self._categorical_conditions = True
        self._condition_dim = int(np.max(np.array(self.train_dataset.targets))) + 1

    def set_criterion(self):
        self._criterion = self._mvae_elbo_loss if 'mvae' in self.parameters['model_name'] else self._elbo_loss

    def _elbo_loss(self, recon_x, x, means, log_var, loss_mask=None, reduce=None, reduction='sum'):
        """problems.py:401-419 (VAE / CVAE) for callers that hold a reconstruction already; inside this
        package the same terms are fused into the step (StepEngine).  reduce=False -> per-sample scores."""
        if reduction != 'sum' or reduce not in (None, False):
            raise NotImplementedError("only reduction='sum' (training) and reduce=False (per-sample scoring)")
        kld = losses.kl_divergence(means, log_var)
        logits = recon_x.view(x.size())
        if reduce is False:
            return losses.bce_with_logits_per_sample(logits, x, loss_mask) + self._kl_weight * kld
        bce = losses.bce_with_logits_sum(logits, x, loss_mask)
        return (bce + self._kl_weight * kld) / x.size(0)

    def _mvae_elbo_loss(self, recon_x, x, means, log_var, loss_mask=None, reduce=None, reduction='sum'):
        """problems.py:421-458: images -> BCE-with-logits (sum), vectors -> MSE (sum) x pose_multiplier."""
        assert len(recon_x) == len(x)
        if reduction != 'sum' or reduce not in (None, False):
            raise NotImplementedError("only reduction='sum' (training) and reduce=False (per-sample scoring)")
        kld = losses.kl_divergence(means, log_var)
        err = 0
        for r, t in zip(recon_x, x):
            if r.dim() > 2:
                r = r.view(t.size())
                e = losses.bce_with_logits_per_sample(r, t, loss_mask) if reduce is False else \
                    losses.bce_with_logits_sum(r, t, loss_mask)
            else:
                if loss_mask is not None:
                    raise ValueError("a (B,3,64,64) loss mask cannot broadcast over (B,7) poses; the reference "
                                     "fails here too (problems.py:446)")
                e = losses.mse_per_sample(r, t, self._pose_multiplier) if reduce is False else \
                    self._pose_multiplier * losses.mse_sum(r, t)
            err = err + e
        if reduce is False:
            return err + self._kl_weight * kld
        return (err + self._kl_weight * kld) / x[0].size(0)

    def _evaluate_mvae_passes(self, x, targets, loss_mask=None, reduce=None, reduction='sum', condition=None):
        """The sub-sampled objective pass by pass through the module-level API, exactly as the reference
        spells it (problems.py:473-546).  Used for the per-sample scoring variant (reduce=False) and as
        an independent cross-check of the fused step; training uses the fused StepEngine."""
        m, up = self._model, self.parameters['use_pose']
        c = dict(condition=condition)
        tv, tt = targets[0], targets[1]
        vj, tj, _, mu, lv = m([x[0], x[1]], **c)
        loss = self._mvae_elbo_loss([vj, tj], [tv, tt], mu, lv, loss_mask, reduce, reduction)
        v1, _, _, mu, lv = m([x[0], None], **c)
        loss = loss + self._mvae_elbo_loss([v1], [tv], mu, lv, loss_mask, reduce, reduction)
        _, t1, _, mu, lv = m([None, x[1]], **c)
        loss = loss + self._mvae_elbo_loss([t1], [tt], mu, lv, loss_mask, reduce, reduction)
        rec = [vj, tj]
        if up:
            vj, tj, pj, mu, lv = m([x[0], x[1]], pose=x[2], **c)
            loss = loss + self._mvae_elbo_loss([vj, tj, pj], [tv, tt, targets[2]], mu, lv, loss_mask, reduce, reduction)
            v2, _, p2, mu, lv = m([x[0], None], pose=x[2], **c)
            loss = loss + self._mvae_elbo_loss([v2, p2], [tv, targets[2]], mu, lv, loss_mask, reduce, reduction)
            _, t2, p3, mu, lv = m([None, x[1]], pose=x[2], **c)
            loss = loss + self._mvae_elbo_loss([t2, p3], [tt, targets[2]], mu, lv, loss_mask, reduce, reduction)
            _, _, p4, mu, lv = m([None, None], pose=x[2], **c)
            loss = loss + self._mvae_elbo_loss([p4], [targets[2]], mu, lv, loss_mask, reduce, reduction)
            rec = [vj, tj, pj]
        return {'recon_x': rec, 'means': mu, 'log_var': lv}, loss

    use_cuda_graph = True  # replay the whole training step as one CUDA graph when the noise source allows it

    def _step_tensors(self, inputs, targets):
        """(x, targets, loss_mask, rename) in the form StepEngine.evaluate takes them."""
        return inputs, inputs, None, None

    def _step_condition(self, inputs):
        """The CVAE condition of this batch (None for un-conditional problems)."""
        if not self._conditional:
            return None
        raise NotImplementedError("class-label (categorical) conditioning is outside the accelerated path")

    def _graphed_step(self, inputs, targets):
        from mmdyn_b200 import noise as _noise
        eng = self._get_engine()
        src = eng._noise()
        if not self.use_cuda_graph or not isinstance(self._optimizer, (fused_optim.FusedAdam, fused_optim.FusedSGD)):
            return None
        cond = self._step_condition(inputs)
        if isinstance(src, _noise.HostNoise):
            if src is not _noise.get_default():
                return None  # a caller-provided host generator must keep its draw order: eager path
            eng.noise_src = src = _noise.DeviceNoise(seed=int(torch.initial_seed() % (1 << 31)))
        x, t, mask, rename = self._step_tensors(inputs, targets)
        first = x[0] if isinstance(x, (list, tuple)) else x
        key = (tuple(first.shape), float(self._kl_weight), mask is not None, float(self._pose_multiplier),
               cond is not None)
        cache = self.__dict__.setdefault("_graph_cache", {})
        g = cache.get(key)
        if g is None:
            if len(cache) > 4:
                cache.clear()
            g = cache[key] = engine.GraphedTrainStep(eng, self._optimizer, x, t, self._kl_weight, loss_mask=mask,
                                                     condition=cond)
        g.load(x, t, mask=mask, condition=cond)
        outputs, loss = g.run()
        if "x" in outputs.get("perf_measure", {}):
            outputs = dict(outputs)
            outputs["perf_measure"] = {rename: outputs["perf_measure"]["x"]} if rename is not None else {}
        return outputs, loss

    def _get_engine(self):
        kind = 'mvae' if ('mvae' in self.parameters['model_name'] and self._cross_modal) else 'vae'
        if self._engine is None or self._engine.model is not self._model:
            self._engine = engine.StepEngine(self._model, kind, use_pose=self.parameters.get('use_pose', False),
                                             pose_multiplier=self._pose_multiplier)
        self._engine.pose_multiplier = float(self._pose_multiplier)
        return self._engine

    def _evaluate_model(self, x, targets, **kwargs):
        if 'mvae' in self.parameters['model_name']:
            return self._evaluate_mvae(x=x, targets=x)
        if self._conditional:
            raise NotImplementedError("class-label (categorical) conditioning of the plain Reconstruction problem "
                                      "is outside the accelerated path; SeqModeling / DynModeling --conditional "
                                      "(shock force) is supported")
        outputs, loss = self._get_engine().evaluate(x, x, self._kl_weight)
        outputs.pop('perf_measure', None)
        return outputs, loss

    def _evaluate_mvae(self, x, targets, loss_mask=None, reduce=None, reduction='sum', condition=None):
        """problems.py:473-546 as one fused step (3 passes, or 7 with --use-pose)."""
        assert isinstance(x, list) and isinstance(targets, list)
        # the reference hands `condition` to every model call; un-conditional encoders / decoders ignore it
        condition = condition if self._conditional else None
        if reduce is not None or reduction != 'sum':
            return self._evaluate_mvae_passes(x, targets, loss_mask, reduce, reduction, condition)
        return self._get_engine().evaluate(x, targets, self._kl_weight, loss_mask=loss_mask, condition=condition)

    def _sample(self, n=50):
        with torch.no_grad():
            if self._conditional:  # problems.py:550-556
                if self._categorical_conditions:
                    raise NotImplementedError("categorical conditions are outside the accelerated path")
                y = torch.rand((n, self._condition_dim)).to(self._device)
                self._img_logger_dict['Samples/latent_space'] = self.apply_sigmoid(self._model.inference(n=n, c=y))
                return
            self._img_logger_dict['Samples/latent_space'] = self.apply_sigmoid(self._model.inference(n=n))

    def _log_train_info(self, inputs, outputs, targets, loss, epoch, perf_measure=None):
        n = len(self.train_loader)
        self._logger_dict['Loss/train_epoch'].append(loss / n)
        self._logger_dict['KL_annealing/train_epoch'].append(self._kl_weight)
        self._img_logger_dict['Input_img/train'] = inputs
        self._img_logger_dict['Output_img/train'] = self.apply_sigmoid(outputs['recon_x'])
        for k, v in (perf_measure or {}).items():
            self._logger_dict['Perf_measure_train/' + k].append(_f(v) / n)

    def _log_test_info(self, inputs, outputs, targets, loss, epoch, perf_measure=None):
        n = len(self.test_loader)
        self._logger_dict['Loss/validation_epoch'].append(loss / n)
        self._img_logger_dict['Input_img/validation'] = inputs
        self._img_logger_dict['Output_img/validation'] = self.apply_sigmoid(outputs['recon_x'])
        for k, v in (perf_measure or {}).items():
            self._logger_dict['Perf_measure_validation/' + k].append(_f(v) / n)
        self._save_if_best(loss, epoch)

    def _save_if_best(self, loss, epoch):
        """Best-validation checkpoint with the reference's keys (problems.py:580-586)."""
        if loss < self._best_loss:
            self.save_checkpoint(loss, epoch)
            self._best_loss = loss

    def _write_images(self, epoch, n_images=120):
        import torchvision
        L, bs = self._seq_length, self.parameters['batchsize']
        sv = 'sv' in self.parameters['dataset_path']
        nrow = L if (L and L > 1 and not sv) else int(math.sqrt(bs))
        if 'modeling' in self.parameters['problem_type'] and not sv:
            n_images = min(bs * (L or 1), n_images)
        else:
            n_images = min(bs, n_images)
        for key, v in self._img_logger_dict.items():
            if isinstance(v, (list, tuple)):
                img = torch.cat((v[0][:n_images], v[1][:n_images]), dim=0)
            else:
                img = v[:n_images]
            self._writer.add_image(key, torchvision.utils.make_grid(img, nrow=nrow), global_step=epoch)

    def apply_sigmoid(self, img):
        """Sigmoid for visualisation only (problems.py:616-626); host-side display path."""
        with torch.no_grad():
            if isinstance(img, (list, tuple)):
                return [torch.sigmoid(t) for t in img]
            return torch.sigmoid(img)


class SeqModeling(Reconstruction, Problem):
    def parse_input(self, data, target):
        """First frame of every sequence -> resting-state target (problems.py:634-673).  Row selection
        `[::L]` happens on the host tensors exactly as in the reference, then the batch is shipped."""
        L, dev, it = self._seq_length, self._device, self.parameters['input_type']
        model_input = target_output = None
        if not isinstance(data, list):
            model_input, target_output = data.to(dev), target.to(dev)
        elif len(data) == 1:
            model_input, target_output = data[0].to(dev), target[0].to(dev)
        elif it in ('visual', 'tactile'):
            k = 0 if it == 'visual' else 1
            model_input, target_output = data[k][::L].to(dev), target[k][::L].to(dev)
        elif it == 'visuotactile':
            model_input = [data[0][::L].to(dev), data[1][::L].to(dev)]
            target_output = [target[0][::L].to(dev), target[1][::L].to(dev)]
        pose = avail = tpose = mask = shock = None
        if isinstance(data, list) and len(data) > 2:
            pose, avail = [data[2][::L].to(dev)], data[3][::L].to(dev)
            tpose, mask = [target[2][::L].to(dev)], target[3][::L].to(dev)
            shock = data[4][::L].to(dev) if len(data) > 4 else None
        return ({'model_input': model_input, 'input_object_pose': pose, 'input_available_modals': avail,
                 'shock': shock},
                {'target_output': target_output, 'target_object_pose': tpose, 'loss_mask': mask})

    def _set_condition_dim(self):
        self._categorical_conditions = False
        try:
            self._condition_dim = len(self.train_dataset.data[0][0][4])
        except Exception:
            self._condition_dim = 0

    def _step_tensors(self, x, targets):
        mask = targets['loss_mask'] if self.parameters['mask_loss'] else None
        if 'mvae' in self.parameters['model_name']:
            if self.parameters['use_pose']:
                if mask is not None:
                    raise ValueError("--mask-loss with --use-pose cannot broadcast a (B,3,64,64) mask over (B,7) "
                                     "poses; the reference fails here too (problems.py:446)")
                return (x['model_input'] + x['input_object_pose'],
                        targets['target_output'] + targets['target_object_pose'], None, None)
            return x['model_input'], targets['target_output'], mask, None
        return x['model_input'], targets['target_output'], mask, self.parameters['input_type']

    def _step_condition(self, x):
        cond = x.get('shock') if self._conditional else None
        if self._conditional and cond is None:
            raise ValueError("--conditional needs the shock force as the 5th data field (datasets.py: data[4])")
        return cond

    def _evaluate_model(self, x, targets, reduction='sum', reduce=None, **kwargs):
        """problems.py:683-716."""
        loss_mask = targets['loss_mask'] if self.parameters['mask_loss'] else None
        x.setdefault('shock', None)
        cond = self._step_condition(x)
        if 'mvae' in self.parameters['model_name']:
            if self.parameters['use_pose']:
                if loss_mask is not None:
                    raise ValueError("--mask-loss with --use-pose cannot broadcast a (B,3,64,64) mask over (B,7) "
                                     "poses; the reference fails here too (problems.py:446)")
                return self._evaluate_mvae(x=x['model_input'] + x['input_object_pose'],
                                           targets=targets['target_output'] + targets['target_object_pose'],
                                           loss_mask=loss_mask, reduce=reduce, reduction=reduction, condition=cond)
            return self._evaluate_mvae(x=x['model_input'], targets=targets['target_output'], loss_mask=loss_mask,
                                       reduce=reduce, reduction=reduction, condition=cond)
        if reduce is not None or reduction != 'sum':
            recon_x, means, log_var = self._model(x['model_input'], cond) if self._conditional \
                else self._model(x['model_input'])
            loss = self._elbo_loss(recon_x, targets['target_output'], means, log_var, loss_mask=loss_mask,
                                   reduce=reduce, reduction=reduction)
            return {'recon_x': recon_x, 'means': means, 'log_var': log_var}, loss
        outputs, loss = self._get_engine().evaluate(x['model_input'], targets['target_output'], self._kl_weight,
                                                    loss_mask=loss_mask, condition=cond)
        outputs['perf_measure'] = {self.parameters['input_type']: outputs['perf_measure']['x']}
        return outputs, loss

    def _log_train_info(self, inputs, outputs, targets, loss, epoch, perf_measure=None, log_pose=False):
        n = len(self.train_loader)
        self._logger_dict['Loss/train_epoch'].append(loss / n)
        self._logger_dict['KL_annealing/train_epoch'].append(self._kl_weight)
        self._img_logger_dict['Input_img/train'] = inputs['model_input']
        self._img_logger_dict['Output_img/train'] = self.apply_sigmoid(outputs['recon_x'])
        self._img_logger_dict['Target_img/train'] = targets['target_output']
        for k, v in (perf_measure or {}).items():
            self._logger_dict['Perf_measure_train/' + k].append(_f(v) / n)

    def _log_test_info(self, inputs, outputs, targets, loss, epoch, perf_measure=None, log_pose=False):
        n = len(self.test_loader)
        self._logger_dict['Loss/validation_epoch'].append(loss / n)
        self._img_logger_dict['Input_img/validation'] = inputs['model_input']
        self._img_logger_dict['Output_img/validation'] = self.apply_sigmoid(outputs['recon_x'])
        self._img_logger_dict['Target_img/validation'] = targets['target_output']
        for k, v in (perf_measure or {}).items():
            self._logger_dict['Perf_measure_validation/' + k].append(_f(v) / n)
        self._save_if_best(loss, epoch)

    def _write_images(self, epoch, n_images=120):
        # recon_x of an MVAE+pose step carries the pose vector as third entry: images only
        for k, v in list(self._img_logger_dict.items()):
            if isinstance(v, (list, tuple)):
                self._img_logger_dict[k] = [t for t in v if t.dim() == 4]
        super()._write_images(epoch, n_images)


class DynModeling(SeqModeling):
    def parse_input(self, data, target):
        """One-step dynamics targets (problems.py:765-803): roll(-1) along the frame axis with every
        last-of-sequence row replaced by the resting-state target; the pose target is the bare roll
        (no end-of-sequence fix-up, :798) — reproduced as is."""
        dev, it = self._device, self.parameters['input_type']
        model_input = target_output = None
        if not isinstance(data, list):
            model_input = data.to(dev)
        else:
            L = self._seq_length

            def shifted(k):
                t = torch.roll(data[k], -1, dims=0).to(dev)
                t[L - 1::L] = target[k][L - 1::L].to(dev)
                return t
            if it in ('visual', 'tactile'):
                k = 0 if it == 'visual' else 1
                model_input, target_output = data[k].to(dev), shifted(k)
            elif it == 'visuotactile':
                model_input = [data[0].to(dev), data[1].to(dev)]
                target_output = [shifted(0), shifted(1)]
        shock = data[4].to(dev) if len(data) > 4 else None
        return ({'model_input': model_input, 'input_object_pose': [data[2].to(dev)],
                 'input_available_modals': data[3].to(dev), 'shock': shock},
                {'target_output': target_output, 'target_object_pose': [torch.roll(data[2], -1, dims=0).to(dev)],
                 'loss_mask': target[3].to(dev)})
