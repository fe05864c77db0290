// LCQProblemBatch.cpp -- see ../include/LCQProblemBatch.hpp.
#include "LCQProblemBatch.hpp"

#include <cstddef>

namespace LCQPow {

namespace {
size_t fieldLen(int k, int n, int c, int p)
{
    switch (k) {
        case LCQP_Q: return (size_t)n * n;
        case LCQP_G: case LCQP_LB: case LCQP_UB: case LCQP_X0: return (size_t)n;
        case LCQP_L: case LCQP_R: return (size_t)p * n;
        case LCQP_LBL: case LCQP_UBL: case LCQP_LBR: case LCQP_UBR: return (size_t)p;
        case LCQP_A: return (size_t)c * n;
        case LCQP_LBA: case LCQP_UBA: return (size_t)c;
        case LCQP_Y0: return (size_t)n + c + 2 * (size_t)p;
    }
    return 0;
}
}  // namespace

LCQProblemBatch::LCQProblemBatch(int _nV, int _nC, int _nComp, int _batch, const std::vector<int>& devices)
    : nV(_nV), nC(_nC), nComp(_nComp), batch(_batch)
{
    std::vector<int> devs = devices.empty() ? std::vector<int>(1, 0) : devices;
    if (batch <= 0) return;
    if ((int)devs.size() > batch) devs.resize((size_t)batch);
    const int G = (int)devs.size();
    // contiguous blocks: shard r owns [r*batch/G, (r+1)*batch/G)
    for (int r = 0; r < G; ++r) {
        Shard s;
        s.device = devs[(size_t)r];
        s.first = (int)((long long)batch * r / G);
        s.count = (int)((long long)batch * (r + 1) / G) - s.first;
        const int rc = lcqp_cuda_create(nV, nC, nComp, s.count, s.device, &s.handle);
        if (rc != LCQP_CUDA_OK) {
            lastError = "lcqp_cuda_create failed on device " + std::to_string(s.device) + " (code " + std::to_string(rc) + ")";
            for (Shard& t : shards) lcqp_cuda_destroy(t.handle);
            shards.clear();
            return;
        }
        lcqp_cuda_set_instance_offset(s.handle, (unsigned long long)s.first);   // same perturbStep draws as the unsharded batch
        shards.push_back(s);
    }
    valid = true;
}

LCQProblemBatch::~LCQProblemBatch()
{
    for (Shard& s : shards) lcqp_cuda_destroy(s.handle);
}

ReturnValue LCQProblemBatch::fail(int code, const Shard& s) const
{
    lastError = std::string(lcqp_cuda_last_error(s.handle)) + " (device " + std::to_string(s.device) + ", code " + std::to_string(code) + ")";
    return (ReturnValue)code;
}

ReturnValue LCQProblemBatch::setOptions(const Options& _options)
{
    if (!valid) return LCQPOBJECT_NOT_SETUP;
    options = _options;
    lcqp_cuda_options o;
    options.toCuda(o);
    for (const Shard& s : shards) {
        const int rc = lcqp_cuda_set_options(s.handle, &o);
        if (rc != LCQP_CUDA_OK) return fail(rc, s);
    }
    return SUCCESSFUL_RETURN;
}

ReturnValue LCQProblemBatch::loadLCQP(unsigned shared_mask, const double* Q, const double* g, const double* L, const double* R,
                                      const double* lbL, const double* ubL, const double* lbR, const double* ubR, const double* A,
                                      const double* lbA, const double* ubA, const double* lb, const double* ub, const double* x0,
                                      const double* y0)
{
    if (!valid) return LCQPOBJECT_NOT_SETUP;
    const double* p[LCQP_NUM_ARRAYS] = {Q, g, L, R, lbL, ubL, lbR, ubR, A, lbA, ubA, lb, ub, x0, y0};
    for (const Shard& s : shards) {
        const double* q[LCQP_NUM_ARRAYS];
        for (int k = 0; k < LCQP_NUM_ARRAYS; ++k) {
            const bool shared = (shared_mask >> k) & 1u;
            q[k] = (p[k] && !shared) ? p[k] + fieldLen(k, nV, nC, nComp) * (size_t)s.first : p[k];
        }
        const int rc = lcqp_cuda_load(s.handle, s.count, shared_mask, q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], q[8], q[9],
                                      q[10], q[11], q[12], q[13], q[14]);
        if (rc != LCQP_CUDA_OK) return fail(rc, s);
    }
    ran = false;
    return SUCCESSFUL_RETURN;
}

ReturnValue LCQProblemBatch::runSolver()
{
    if (!valid) return LCQPOBJECT_NOT_SETUP;
    for (const Shard& s : shards) {   // asynchronous launches: the shards run concurrently
        const int rc = lcqp_cuda_run(s.handle, nullptr);
        if (rc != LCQP_CUDA_OK) return fail(rc, s);
    }
    for (const Shard& s : shards) {
        const int rc = lcqp_cuda_synchronize(s.handle);
        if (rc != LCQP_CUDA_OK) return fail(rc, s);
    }
    ran = true;
    return SUCCESSFUL_RETURN;
}

ReturnValue LCQProblemBatch::getPrimalSolution(double* x) const
{
    if (!valid || !ran) return LCQPOBJECT_NOT_SETUP;
    for (const Shard& s : shards) {   // host-side gather at the shard's offset
        const int rc = lcqp_cuda_get_primal(s.handle, x + (size_t)s.first * nV);
        if (rc != LCQP_CUDA_OK) return fail(rc, s);
    }
    return SUCCESSFUL_RETURN;
}

ReturnValue LCQProblemBatch::getDualSolution(double* y) const
{
    if (!valid || !ran) return LCQPOBJECT_NOT_SETUP;
    const size_t nD = (size_t)nV + nC + 2 * (size_t)nComp;
    for (const Shard& s : shards) {
        const int rc = lcqp_cuda_get_dual(s.handle, y + (size_t)s.first * nD);
        if (rc != LCQP_CUDA_OK) return fail(rc, s);
    }
    return SUCCESSFUL_RETURN;
}

ReturnValue LCQProblemBatch::getRawStatistics(std::vector<lcqp_cuda_stats>& out) const
{
    if (!valid || !ran) return LCQPOBJECT_NOT_SETUP;
    out.resize((size_t)batch);
    for (const Shard& s : shards) {
        const int rc = lcqp_cuda_get_stats(s.handle, out.data() + s.first);
        if (rc != LCQP_CUDA_OK) return fail(rc, s);
    }
    return SUCCESSFUL_RETURN;
}

ReturnValue LCQProblemBatch::getOutputStatistics(std::vector<OutputStatistics>& out) const
{
    std::vector<lcqp_cuda_stats> raw;
    const ReturnValue r = getRawStatistics(raw);
    if (r != SUCCESSFUL_RETURN) return r;
    out.resize(raw.size());
    for (size_t b = 0; b < raw.size(); ++b) out[b].fromCuda(raw[b]);
    return SUCCESSFUL_RETURN;
}

std::vector<int> LCQProblemBatch::getReturnValues() const
{
    std::vector<lcqp_cuda_stats> raw;
    std::vector<int> ret;
    if (getRawStatistics(raw) != SUCCESSFUL_RETURN) return ret;
    ret.resize(raw.size());
    for (size_t b = 0; b < raw.size(); ++b) ret[b] = raw[b].ret;
    return ret;
}

int LCQProblemBatch::getNumberOfDuals() const
{
    return shards.empty() ? 0 : lcqp_cuda_num_duals(shards[0].handle);
}

long long LCQProblemBatch::getLaunchCount() const
{
    long long n = 0;
    for (const Shard& s : shards) n += lcqp_cuda_launch_count(s.handle);
    return n;
}

}  // namespace LCQPow
