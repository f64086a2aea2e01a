"""Small model builders shared by the golden-vector generator and the tests (TEST INFRASTRUCTURE)."""

import torch
from torch import nn


def mlp_c1(classes: int = 10, width: int = 64) -> nn.Sequential:
    """BASELINE.json configs[0]: D=64, 4 Linear + ReLU, 13,130 parameters."""
    return nn.Sequential(
        nn.Linear(width, width), nn.ReLU(), nn.Linear(width, width), nn.ReLU(),
        nn.Linear(width, width), nn.ReLU(), nn.Linear(width, classes),
    )


def mlp_smooth(width: int = 16, classes: int = 6) -> nn.Sequential:
    """Linear-Sigmoid-Linear-Tanh-Linear: smooth activations, so the Hessian has second-order terms."""
    return nn.Sequential(nn.Linear(width, width), nn.Sigmoid(), nn.Linear(width, width), nn.Tanh(),
                         nn.Linear(width, classes))


class BasicBlock(nn.Module):
    """torchvision-style residual block (conv-bn-relu-conv-bn + shortcut, relu)."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False),
                                            nn.BatchNorm2d(cout))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + idt)


class MiniResNet(nn.Module):
    """A ResNet-18-shaped network at toy size: 7x7/2 stem, maxpool(3,2,1), two stages."""

    def __init__(self, width=8, classes=10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = nn.Sequential(BasicBlock(width, width, 1))
        self.layer2 = nn.Sequential(BasicBlock(width, 2 * width, 2))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(2 * width, classes)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer2(self.layer1(x))
        return self.fc(torch.flatten(self.avgpool(x), 1))


def randomize_bn_(model: nn.Module, gen: torch.Generator) -> None:
    """Give eval-mode BatchNorm non-trivial running statistics and affine parameters."""
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.running_mean.copy_(0.3 * torch.randn(m.num_features, generator=gen))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=gen))
            with torch.no_grad():
                m.weight.copy_(0.5 + torch.rand(m.num_features, generator=gen))
                m.bias.copy_(0.2 * torch.randn(m.num_features, generator=gen))


class ConvNetBias(nn.Module):
    """Plain CNN with conv biases, stride/padding variety and flatten->Linear (KFAC cases)."""

    def __init__(self, classes=5):
        super().__init__()
        self.c1 = nn.Conv2d(3, 6, 3, 1, 1)
        self.c2 = nn.Conv2d(6, 8, 3, 2, 0)
        self.fc = nn.Linear(8 * 3 * 3, classes)

    def forward(self, x):
        x = torch.relu(self.c1(x))
        x = torch.relu(self.c2(x))
        return self.fc(torch.flatten(x, 1))


def mlp_ln_gelu(width: int = 16, classes: int = 6) -> nn.Sequential:
    """Linear-LayerNorm-GELU-Linear-LayerNorm-ReLU-Linear (north_star layer set + GELU)."""
    return nn.Sequential(nn.Linear(width, 20), nn.LayerNorm(20), nn.GELU(), nn.Linear(20, 12), nn.LayerNorm(12),
                         nn.ReLU(), nn.Linear(12, classes))


def mlp_gelu(width: int = 16, classes: int = 6) -> nn.Sequential:
    """Linear-GELU-Linear-GELU-Linear: smooth activation, the Hessian has second-order terms."""
    return nn.Sequential(nn.Linear(width, width), nn.GELU(), nn.Linear(width, width), nn.GELU(),
                         nn.Linear(width, classes))


class TokenMLP(nn.Module):
    """Token sequence [B, T, D] -> pre-norm MLP block with a residual, LayerNorm, mean over tokens, linear head: the
    Linear / LayerNorm / GELU / residual pattern of a transformer block without the attention."""

    def __init__(self, dim: int = 12, hidden: int = 20, classes: int = 5):
        super().__init__()
        self.ln1, self.fc1, self.fc2 = nn.LayerNorm(dim), nn.Linear(dim, hidden), nn.Linear(hidden, dim)
        self.ln2, self.head = nn.LayerNorm(dim), nn.Linear(dim, classes)

    def forward(self, x):
        x = x + self.fc2(torch.nn.functional.gelu(self.fc1(self.ln1(x))))
        return self.head(self.ln2(x).mean(1))


class TiedNet(nn.Module):
    """Weight tying: one Conv2d and one Linear applied twice each (reference io_collector/groups.py:123-168 concatenates
    the usages of a tied parameter along the weight-sharing axis)."""

    def __init__(self, classes: int = 5):
        super().__init__()
        self.stem = nn.Conv2d(3, 4, 3, padding=1)
        self.conv = nn.Conv2d(4, 4, 3, padding=1)
        self.fc = nn.Linear(4, 4)
        self.head = nn.Linear(4, classes)

    def forward(self, x):
        x = torch.tanh(self.stem(x))
        x = torch.nn.functional.max_pool2d(torch.tanh(self.conv(x)), 2)   # first usage on 8 x 8, second on 4 x 4
        x = torch.tanh(self.conv(x))
        x = torch.nn.functional.adaptive_avg_pool2d(x, 1).flatten(1)
        x = torch.tanh(self.fc(torch.tanh(self.fc(x))))
        return self.head(x)


class TransformerBlock(nn.Module):
    """Pre-norm transformer encoder block on token sequences [B, T, D] (LayerNorm, nn.MultiheadAttention with
    batch_first, residuals, Linear-GELU-Linear), mean over tokens, linear head: the ViT encoder layer without the
    class token / position embedding."""

    def __init__(self, dim: int = 16, heads: int = 2, hidden: int = 32, classes: int = 5, layers: int = 1,
                 sharpen: bool = False):
        super().__init__()
        self.blocks = nn.ModuleList()
        for _ in range(layers):
            blk = nn.Module()
            blk.ln1, blk.attn, blk.ln2 = nn.LayerNorm(dim), nn.MultiheadAttention(dim, heads, batch_first=True), nn.LayerNorm(dim)
            blk.fc1, blk.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)
            self.blocks.append(blk)
        self.ln, self.head = nn.LayerNorm(dim), nn.Linear(dim, classes)
        if sharpen:  # the default initialisation leaves the attention nearly uniform
            with torch.no_grad():
                for blk in self.blocks:
                    blk.attn.in_proj_weight.mul_(3.0)
                    blk.attn.in_proj_bias.normal_(0.0, 0.3)

    def forward(self, x):
        for blk in self.blocks:
            y = blk.ln1(x)
            x = x + blk.attn(y, y, y, need_weights=False)[0]
            x = x + blk.fc2(torch.nn.functional.gelu(blk.fc1(blk.ln2(x))))
        return self.head(self.ln(x).mean(1))


def mini_vit(classes: int = 5, layers: int = 2):
    """torchvision's VisionTransformer at toy size (32 x 32 images, 8 x 8 patches, 2 heads, width 32) with generic values
    for the parameters torchvision zero-initialises (classification head) or leaves tiny (class token, positions)."""
    from torchvision.models.vision_transformer import VisionTransformer

    model = VisionTransformer(image_size=32, patch_size=8, num_layers=layers, num_heads=2, hidden_dim=32, mlp_dim=64,
                              num_classes=classes)
    with torch.no_grad():
        model.heads.head.weight.normal_(0.0, 0.3)
        model.heads.head.bias.normal_(0.0, 0.1)
        model.class_token.normal_(0.0, 0.5)
        model.encoder.pos_embedding.normal_(0.0, 0.5)
        for blk in model.encoder.layers:
            blk.self_attention.in_proj_weight.mul_(3.0)
            blk.self_attention.in_proj_bias.normal_(0.0, 0.3)
    return model
