// engine_internal.h — C++-side accessors shared by engine.cu and symexp.cpp (not part of the C-ABI).
#pragma once
#include "../../include/symgpu.h"
#include "flatnet.h"
const symb200::FlatNet& symgpu_internal_net(symgpu_ctx* c);
// queues Reseau::CreateVehicle (reseau.cpp:13351) for the next step; returns the new vehicle id
int symgpu_internal_create_vehicle(symgpu_ctx* c, int cls, int entry, int dest, int lane, double dbt);
// vehicle type of a vehicle id (known to the host since the vehicle was created); -1 unknown
int symgpu_internal_class_of(symgpu_ctx* c, int id);
