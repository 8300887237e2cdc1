"""Drop-in ConvolutionalEncoder, DeconvolutionalDecoder and ConvolutionalVQVAE.

Same constructors, forward signatures, attribute / state_dict names and construction order (hence identical
initial weights under the same torch seed) as the reference's src/models/convolutional_encoder.py:36-146,
deconvolutional_decoder.py:38-137 and convolutional_vq_vae.py:38-139.
"""
import torch
import torch.nn as nn

from . import functional as F
from . import ops
from .modules import Conv1DBuilder, Conv1d, ConvTranspose1DBuilder, Jitter, ResidualStack
from .vector_quantizer import VectorQuantizer, VectorQuantizerEMA


class ConvolutionalEncoder(nn.Module):

    def __init__(self, in_channels, num_hiddens, num_residual_layers, num_residual_hiddens, use_kaiming_normal,
                 input_features_type, features_filters, sampling_rate, device, verbose=False):
        super(ConvolutionalEncoder, self).__init__()
        # note: like the reference (convolutional_encoder.py:49-50) the first conv takes `features_filters` input
        # channels; `in_channels` is accepted and ignored
        self._conv_1 = Conv1DBuilder.build(features_filters, num_hiddens, 3, use_kaiming_normal=use_kaiming_normal,
                                           padding=1)
        self._conv_2 = Conv1DBuilder.build(num_hiddens, num_hiddens, 3, use_kaiming_normal=use_kaiming_normal,
                                           padding=1)
        self._conv_3 = Conv1DBuilder.build(num_hiddens, num_hiddens, 4, stride=2,
                                           use_kaiming_normal=use_kaiming_normal, padding=2)
        self._conv_4 = Conv1DBuilder.build(num_hiddens, num_hiddens, 3, use_kaiming_normal=use_kaiming_normal,
                                           padding=1)
        self._conv_5 = Conv1DBuilder.build(num_hiddens, num_hiddens, 3, use_kaiming_normal=use_kaiming_normal,
                                           padding=1)
        self._residual_stack = ResidualStack(in_channels=num_hiddens, num_hiddens=num_hiddens,
                                             num_residual_layers=num_residual_layers,
                                             num_residual_hiddens=num_residual_hiddens,
                                             use_kaiming_normal=use_kaiming_normal)
        self._input_features_type = input_features_type
        self._features_filters = features_filters
        self._sampling_rate = sampling_rate
        self._device = device
        self._verbose = verbose

    def forward(self, inputs):
        x_conv_1 = self._conv_1(inputs, relu=True)
        x = F.add(self._conv_2(x_conv_1, relu=True), x_conv_1)
        x_conv_3 = self._conv_3(x, relu=True)
        x_conv_4 = F.add(self._conv_4(x_conv_3, relu=True), x_conv_3)
        x_conv_5 = F.add(self._conv_5(x_conv_4, relu=True), x_conv_4)
        return F.add(self._residual_stack(x_conv_5), x_conv_5)


class DeconvolutionalDecoder(nn.Module):

    def __init__(self, in_channels, out_channels, num_hiddens, num_residual_layers, num_residual_hiddens,
                 use_kaiming_normal, use_jitter, jitter_probability, use_speaker_conditioning, device, verbose=False):
        super(DeconvolutionalDecoder, self).__init__()
        self._use_jitter = use_jitter
        self._use_speaker_conditioning = use_speaker_conditioning
        self._device = device
        self._verbose = verbose
        if self._use_jitter:
            self._jitter = Jitter(jitter_probability)
        # deconvolutional_decoder.py:56: the speaker features add 40 input channels ("FIXME hardcoded" there)
        in_channels = in_channels + 40 if self._use_speaker_conditioning else in_channels
        self._conv_1 = Conv1DBuilder.build(in_channels, num_hiddens, 3, padding=1,
                                           use_kaiming_normal=use_kaiming_normal)
        self._upsample = nn.Upsample(scale_factor=2)   # kept for attribute parity; the kernel is vqs_upsample2_fwd
        self._residual_stack = ResidualStack(in_channels=num_hiddens, num_hiddens=num_hiddens,
                                             num_residual_layers=num_residual_layers,
                                             num_residual_hiddens=num_residual_hiddens,
                                             use_kaiming_normal=use_kaiming_normal)
        self._conv_trans_1 = ConvTranspose1DBuilder.build(num_hiddens, num_hiddens, 3, padding=1,
                                                          use_kaiming_normal=use_kaiming_normal)
        self._conv_trans_2 = ConvTranspose1DBuilder.build(num_hiddens, num_hiddens, 3, padding=0,
                                                          use_kaiming_normal=use_kaiming_normal)
        self._conv_trans_3 = ConvTranspose1DBuilder.build(num_hiddens, out_channels, 2, padding=0,
                                                          use_kaiming_normal=use_kaiming_normal)

    def forward(self, inputs, speaker_dic, speaker_id, out_len=None):
        x = inputs
        if self._use_jitter and self.training:
            x = self._jitter(x)
        if self._use_speaker_conditioning:
            # deconvolutional_decoder.py:108-111 / global_conditioning.py:34-57: a FRESH nn.Embedding(len(speaker_dic), 40)
            # drawn N(0, 0.1) on every call (host RNG; same constructor + normal_ sequence, so the same seed gives the same
            # features), looked up per utterance and repeated over time.  Host-side plumbing: no arithmetic.
            B, _, T = x.shape
            emb = nn.Embedding(len(speaker_dic), 40, padding_idx=None)
            emb.weight.data.normal_(0, 0.1)
            emb = emb.to(x.device)
            rows = emb.weight.data[speaker_id.to(x.device).view(B).long()]                # (B, 40): a row lookup
            x = F.concat_channels(x, rows)                                               # vqs_concat_channels
        x = self._conv_1(x)
        x = F.upsample2(x)
        x = self._residual_stack(x)
        x = self._conv_trans_1(x, relu=True)
        x = self._conv_trans_2(x, relu=True)
        return self._conv_trans_3(x, out_len=out_len)


class ConvolutionalVQVAE(nn.Module):

    def __init__(self, configuration, device):
        super(ConvolutionalVQVAE, self).__init__()
        self._output_features_filters = configuration['output_features_filters'] * 3 \
            if configuration['augment_output_features'] else configuration['output_features_filters']
        self._output_features_dim = configuration['output_features_dim']
        self._verbose = configuration['verbose']
        self._encoder = ConvolutionalEncoder(
            in_channels=configuration['input_features_dim'],
            num_hiddens=configuration['num_hiddens'],
            num_residual_layers=configuration['num_residual_layers'],
            num_residual_hiddens=configuration['num_hiddens'],
            use_kaiming_normal=configuration['use_kaiming_normal'],
            input_features_type=configuration['input_features_type'],
            features_filters=configuration['input_features_filters'] * 3
            if configuration['augment_input_features'] else configuration['input_features_filters'],
            sampling_rate=configuration['sampling_rate'],
            device=device,
            verbose=self._verbose)
        self._pre_vq_conv = Conv1d(in_channels=configuration['num_hiddens'],
                                   out_channels=configuration['embedding_dim'], kernel_size=3, padding=1)
        if configuration['decay'] > 0.0:
            self._vq = VectorQuantizerEMA(num_embeddings=configuration['num_embeddings'],
                                          embedding_dim=configuration['embedding_dim'],
                                          commitment_cost=configuration['commitment_cost'],
                                          decay=configuration['decay'], device=device)
        else:
            self._vq = VectorQuantizer(num_embeddings=configuration['num_embeddings'],
                                       embedding_dim=configuration['embedding_dim'],
                                       commitment_cost=configuration['commitment_cost'], device=device)
        self._vq.materialize_outputs = False      # the model discards encodings / distances (convolutional_vq_vae.py:128)
        self._decoder = DeconvolutionalDecoder(
            in_channels=configuration['embedding_dim'],
            out_channels=self._output_features_filters,
            num_hiddens=configuration['num_hiddens'],
            num_residual_layers=configuration['num_residual_layers'],
            num_residual_hiddens=configuration['residual_channels'],
            use_kaiming_normal=configuration['use_kaiming_normal'],
            use_jitter=configuration['use_jitter'],
            jitter_probability=configuration['jitter_probability'],
            use_speaker_conditioning=configuration['use_speaker_conditioning'],
            device=device,
            verbose=self._verbose)
        self._device = device
        self._record_codebook_stats = configuration['record_codebook_stats']

    @property
    def vq(self):
        return self._vq

    @property
    def pre_vq_conv(self):
        return self._pre_vq_conv

    @property
    def encoder(self):
        return self._encoder

    @property
    def decoder(self):
        return self._decoder

    def forward(self, x, speaker_dic, speaker_id):
        if not x.is_cuda:
            raise RuntimeError('ConvolutionalVQVAE: inputs must be on a CUDA device (no CPU fallback)')
        x = ops.blc_to_ncl(x.float().contiguous())      # .permute(0, 2, 1).contiguous().float()  (vq_vae.py:118)
        z = self._encoder(x)
        z = self._pre_vq_conv(z)
        vq_loss, quantized, perplexity, _, _, encoding_indices, losses, _, _, _, concatenated_quantized = \
            self._vq(z, record_codebook_stats=self._record_codebook_stats)
        # the reference decodes 2*T_q + 3 positions and then drops the tail (vq_vae.py:133-137); computing only the
        # kept positions gives the same tensor and the same gradients
        reconstructed_x = self._decoder(quantized, speaker_dic, speaker_id, out_len=x.size(2))
        reconstructed_x = reconstructed_x.view(-1, self._output_features_filters, x.size(2))
        return reconstructed_x, vq_loss, losses, perplexity, encoding_indices, concatenated_quantized
