// check_epoch.h -- instrumentation injected (nvcc -include) into BOTH generated programs of a pair by
// host/codegen/build_models.sh: at epochs 1, 2 and 5 the training loop prints
//     CHECK <epoch> <sum|prediction|> <loss> <sum|grad| of every parameter, net->named_parameters() order>
// after loss.backward() + optimizer.step() (the gradients are still in place; the next epoch zeroes them).
// Epoch 1 pins the forward pass; its gradients pin the backward kernels (aggregation over the slot 2li+1
// graph, SDDMM, softmax / edge-sum backward); epochs 2 and 5 pin Adam + repeated steps
// (reference src/codegen/common.h:835-977, 1476-1477, 1523-1528).  Sums are taken in fp64.
#pragma once
#include <torch/torch.h>

#include <iomanip>
#include <iostream>

template <class Net>
inline void gala_check_epoch(size_t epoch, const torch::Tensor& prediction, const torch::Tensor& loss, Net& net) {
    if (epoch != 1 && epoch != 2 && epoch != 5) return;
    torch::NoGradGuard ng;
    std::cout << "CHECK " << epoch << std::setprecision(12) << " "
              << prediction.to(torch::kFloat64).abs().sum().item<double>() << " " << loss.item<float>();
    for (auto& kv : net->named_parameters()) {
        const torch::Tensor& g = kv.value().grad();
        std::cout << " " << kv.key() << "="
                  << (g.defined() ? g.to(torch::kFloat64).abs().sum().item<double>() : -1.0);
    }
    std::cout << std::endl;
}
