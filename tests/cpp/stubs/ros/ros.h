// stand-in for <ros/ros.h>: the global parameter server (ros::param::get, typed like roscpp's) and the two macros
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <map>
#include <string>
namespace ros {
namespace param {
inline std::map<std::string, double>& stub_store() { static std::map<std::string, double> m; return m; }
inline std::map<std::string, char>& stub_types() { static std::map<std::string, char> m; return m; }  // 'd' double, 'i' int, 'b' bool
inline void set(const std::string& k, double v) { stub_store()[k] = v; stub_types()[k] = 'd'; }
inline void set(const std::string& k, int v) { stub_store()[k] = v; stub_types()[k] = 'i'; }
inline void set(const std::string& k, bool v) { stub_store()[k] = v ? 1 : 0; stub_types()[k] = 'b'; }
inline bool stub_get(const std::string& k, char type, double& v) {
  auto it = stub_store().find(k);
  if (it == stub_store().end()) return false;
  const char t = stub_types()[k];
  if (t != type && !(type == 'd' && t == 'i')) return false;  // roscpp converts int -> double, nothing else
  v = it->second;
  return true;
}
inline bool get(const std::string& k, double& v) { return stub_get(k, 'd', v); }
inline bool get(const std::string& k, float& v) { double d; if (!stub_get(k, 'd', d)) return false; v = (float)d; return true; }
inline bool get(const std::string& k, int& v) { double d; if (!stub_get(k, 'i', d)) return false; v = (int)d; return true; }
inline bool get(const std::string& k, bool& v) { double d; if (!stub_get(k, 'b', d)) return false; v = d != 0; return true; }
}  // namespace param
}  // namespace ros
#define ROS_BREAK() abort()
#define ROS_INFO(...) do { printf(__VA_ARGS__); printf("\n"); } while (0)
