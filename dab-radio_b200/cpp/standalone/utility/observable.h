// Stand-in for the reference's src/utility/observable.h (same interface: Attach / Notify) for builds outside its tree.
#pragma once
#include <functional>
#include <utility>
#include <vector>
template <typename... Args>
class Observable {
public:
    using Callback = std::function<void(Args...)>;
    void Attach(const Callback& cb) { m_callbacks.push_back(cb); }
    void Notify(Args... args) {
        for (auto& cb : m_callbacks) cb(args...);
    }
private:
    std::vector<Callback> m_callbacks;
};
